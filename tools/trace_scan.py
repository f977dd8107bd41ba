"""Dumps the pipeline trace of the one-pass scan (MQ_SCAN_TRACE) as per-event mean offsets in SM clocks."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from most_queue_b200 import _lib

nr = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
ctx = _lib.default_context()
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.empty(nr + 1, device="cuda", dtype=torch.float64).exponential_(1.0, generator=g)
S = torch.empty(nr, device="cuda", dtype=torch.float64).exponential_(1.0 / 0.7, generator=g)
torch.cuda.synchronize()
ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
os.environ["MQ_SCAN_TRACE"] = "/tmp/scan_trace.bin"
ctx.scan_apply(nr, 0.0, A=A.data_ptr(), S=S.data_ptr(), device_ptrs=True)
os.environ.pop("MQ_SCAN_TRACE")
tr = np.fromfile("/tmp/scan_trace.bin", dtype=np.int64).reshape(-1, 64, 24)
names = ["f_tma_issue", "w_tma_issue", "fold_done", "sum_published", "lb_start", "lb_scanned", "incf_published",
         "prefix_ready", "inc_published", "walk_start", "walk_done"]
for cta in (0, 73, 147):
    if cta >= tr.shape[0]:
        continue
    t = tr[cta]
    print(f"CTA {cta}: per-tile interval (walk_done[k] - walk_done[k-1]) median {np.median(np.diff(t[8:60, 10])):.0f} clk")
    for k in (10, 11, 12, 30):
        base = t[k, 0]
        print(f"  tile {k}: " + " ".join(f"{n}={t[k, i] - base}" for i, n in enumerate(names)) + f" depth={t[k, 11]}")
d = tr[:, 8:60, :]
for a, b in [(0, 2), (2, 3), (3, 4), (4, 5), (5, 6), (3, 7), (3, 1), (1, 9), (7, 9), (9, 10), (0, 10)]:
    print(f"{names[a]:>15} -> {names[b]:<15} median {np.median(d[:, :, b] - d[:, :, a]):8.0f} clk   p90 {np.percentile(d[:, :, b] - d[:, :, a], 90):8.0f}")
print("look-back A depth: median", np.median(d[:, :, 11]), "max", d[:, :, 11].max(), " look-back B depth (-1 = chunk start): median", np.median(d[:, :, 12]), "max", d[:, :, 12].max(), "share at chunk start", (d[:, :, 12] < 0).mean())
print("per-tile interval over all CTAs: median", np.median(np.diff(d[:, :, 10], axis=1)))

# wave across CTAs at one iteration (global ns clock)
k = 20
g0 = tr[:, k, 16].min()
np.set_printoptions(linewidth=250)
print("iteration", k, ": per CTA (every 4th): sum_pub, incf_pub, step2_start, incf_waits_done, carry_wait_done, cex_pub(0 if not first), lbB_gathered, prefix_ready, tma_issue, walk_done  [ns after the first sum_pub]")
print("columns: sum_pub incf_pub | step2_start early_waits_done ks_done late_arrived cex_pub | lbB_gathered prefix_ready | f_tma_issue walk_done")
for b in list(range(8, 36)) + list(range(100, 120)):
    row = tr[b, k]
    print(b, (b + 148 * k) % 16, [int(row[i] - g0) if row[i] else 0 for i in (16, 17, 13, 14, 15, 23, 18, 20, 19, 21, 22)])

print("look-back A of first-of-chunk tiles, SM clocks after lb_start: scanned, incf_published, step2_start, early_waits_done, ks_done, late_arrived(lane l-1)")
first = tr[:, 8:60, 13] != 0
for i, n in [(5, "scanned"), (6, "incf_pub"), (13, "step2_start"), (14, "early_waits_done"), (15, "ks_done"), (23, "late_arrived")]:
    dd = (tr[:, 8:60, i] - tr[:, 8:60, 4])[first & (tr[:, 8:60, i] != 0)]
    print(f"  {n:>18}: median {np.median(dd):8.0f}  p90 {np.percentile(dd, 90):8.0f}  n={dd.size}")
