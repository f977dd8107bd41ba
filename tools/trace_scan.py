"""Dumps the pipeline trace of the one-pass scan (MQ_SCAN_TRACE) as per-event mean offsets in SM clocks.

Events per (CTA, tile k):  0 producer: stage free, TMA issued   9 compute: tile in shared memory (fold starts)
  2 scan done   3 summary published   10 provisional walk done (stage released)   4 look-back A starts   5 A gathered
  6 incf published   7 B: exclusive prefix known   8 inc published   13 segments repaired   14 repair done
  11 / 12 look-back depth of A / B   16.. global ns clocks
"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib

nr = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
rho = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
ctx = _lib.default_context()
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.empty(nr + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
S = torch.empty(nr, device="cuda", dtype=torch.float64).exponential_(1.0 / rho, generator=g)
torch.cuda.synchronize()
ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
os.environ["MQ_SCAN_TRACE"] = "/tmp/scan_trace.bin"
ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
os.environ.pop("MQ_SCAN_TRACE")
tr = np.fromfile("/tmp/scan_trace.bin", dtype=np.int64).reshape(-1, 64, 24)
names = {0: "tma_issue", 9: "tile_in_smem", 2: "scan_done", 3: "sum_published", 10: "walk_done", 4: "lbA_start", 5: "lbA_gathered",
         6: "incf_published", 7: "prefix_known", 8: "inc_published", 14: "repair_done"}
for cta in (0, 73, 147):
    if cta >= tr.shape[0]:
        continue
    t = tr[cta]
    print(f"CTA {cta}: per-tile interval (walk_done[k] - walk_done[k-1]) median {np.median(np.diff(t[8:60, 10])):.0f} clk")
    for k in (10, 11, 12, 30):
        base = t[k, 0]
        print(f"  tile {k}: " + " ".join(f"{n}={t[k, i] - base}" for i, n in names.items()) + f" repaired={t[k, 13]}")
d = tr[:, 8:60, :]
for a, b in [(0, 9), (9, 2), (2, 3), (3, 10), (9, 10), (3, 4), (4, 5), (5, 6), (3, 7), (7, 14), (3, 14)]:
    x = d[:, :, b] - d[:, :, a]
    print(f"{names[a]:>15} -> {names[b]:<15} median {np.median(x):8.0f} clk   p90 {np.percentile(x, 90):8.0f}")
# how long the compute group sat between finishing its previous tile and getting this one (group = k % 2)
gap = d[:, 2:, 9] - d[:, :-2, 10]
print(f"group idle between its tiles (tile_in_smem[k] - walk_done[k-2]): median {np.median(gap):.0f} p90 {np.percentile(gap, 90):.0f}")
# stage turnaround: the producer issues tile k when tile k-3 released the stage
rel = d[:, 3:, 0] - d[:, :-3, 10]
print(f"producer issue after the stage was released (tma_issue[k] - walk_done[k-3]): median {np.median(rel):.0f}")
print("segments repaired per tile: mean", d[:, :, 13].mean(), "max", d[:, :, 13].max())
print("look-back depth A: median", np.median(d[:, :, 11]), " B (-1 = chunk start): median", np.median(d[:, :, 12]))
print("per-tile interval over all CTAs: median", np.median(np.diff(d[:, :, 10], axis=1)))
