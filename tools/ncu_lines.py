"""Per-source-line issue counts of one profiled kernel.

Joins the SASS rows of an `ncu --set full --import-source on` report (executed warp instructions and
thread instructions per SASS address) with the line table of the cubin the kernel came from
(`nvdisasm -g`), and prints the source lines that issue the most, with the average number of active
lanes on each.  The object file must be the build that was profiled.

    python tools/ncu_lines.py gpurun_out/x.ncu-rep most_queue_b200/csrc/_build/k4v2_prio.o k4v2_prioILi4ELi2ELb0 [top] [--by-samples]

The second column is the line's share of the warp-state samples (where warps WAIT), which is what matters for a
latency-bound kernel; --by-samples sorts by it.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    return [(int(r[0], 16) - base, r[ix["Source"]].strip(), int(r[ix["Instructions Executed"]]),
             int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]] or 0)) for r in data]


def line_table(obj, mangled_part):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    table = {}
    inside = False
    cur = ("?", 0)
    for ln in txt.splitlines():
        if ln.startswith("\t.section\t.text."):
            inside = mangled_part in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, obj, name = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else 40
    rows = sass_rows(rep)
    table = line_table(obj, name)
    if len(table) != len(rows):
        print(f"warning: {len(rows)} SASS rows in the report, {len(table)} in the object (different builds?)")
    per = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = sum(r[2] for r in rows)
    tot_s = max(1, sum(r[4] for r in rows))
    for off, _, ie, te, smp in rows:
        k = table.get(off, ("?", 0))
        per[k][0] += ie
        per[k][1] += te
        per[k][2] += 1
        per[k][3] += smp
    by_samples = "--by-samples" in sys.argv
    src_cache = {}

    def text(k):
        f, l = k
        if f not in src_cache:
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "most_queue_b200", "csrc", f)
            src_cache[f] = open(p).read().splitlines() if os.path.exists(p) else []
        s = src_cache[f]
        return s[l - 1].strip()[:90] if 0 < l <= len(s) else ""

    print(f"total warp instructions {tot:.4g}, thread instructions {sum(r[3] for r in rows):.4g}")
    for k, (ie, te, n, smp) in sorted(per.items(), key=lambda kv: -(kv[1][3] if by_samples else kv[1][0]))[:top]:
        print(f"{ie / tot * 100:5.1f}%  stall-samples {smp / tot_s * 100:5.1f}%  lanes {te / max(ie, 1):4.1f}  sass {n:4d}  "
              f"{k[0]}:{k[1]:<4d} {text(k)}")


if __name__ == "__main__":
    main()
