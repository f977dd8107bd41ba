"""Builds most_queue_b200/libmqb200.so in-tree with nvcc for sm_100a.

The kernels are instantiated once per server-slot count in separate translation units so the
build parallelises; objects land in most_queue_b200/csrc/_build/ and are rebuilt only when a
source they depend on is newer.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libmqb200.so")

SLOT_COUNTS = [1, 2, 3, 4, 5, 6, 7, 8, 16, 32]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "186",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libmqb200.so cannot be built")


def _units():
    units = [("mq_api.cu", [], "mq_api.o"), ("reduce.cu", [], "reduce.o")]
    for n in SLOT_COUNTS:
        units.append(("k2_inst.cu", [f"-DMQ_CT={n}"], f"k2_ct{n}.o"))
        units.append(("k1_inst.cu", [f"-DMQ_CT={n}"], f"k1_ct{n}.o"))
    for extra in ("mq_api_prio.cu", "mq_api_net.cu", "mq_api_scan.cu", "mq_api_tandem.cu", "mq_api_prionet.cu", "mq_api_vacation.cu", "mq_api_negatives.cu", "mq_api_size.cu", "k3_lindley.cu", "k3_onepass.cu", "k1b_fifo_jobs.cu",
                  "k4_prio.cu", "k4v2_prio.cu", "k5_net.cu", "k5v2_net.cu", "k6_tandem.cu", "k6v2_tandem.cu", "k7_prionet.cu", "k7v2_prionet.cu", "k8_vacation.cu", "k9_negatives.cu", "k10_size.cu"):
        if os.path.exists(os.path.join(CSRC, extra)):
            units.append((extra, [], extra[:-3] + ".o"))
    return units


def _newest_header() -> float:
    t = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = _newest_header()
    todo = []
    objs = []
    for src, defs, obj in _units():
        s, o = os.path.join(CSRC, src), os.path.join(OBJ, obj)
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_t):
            todo.append([nvcc, *NVCC_FLAGS, *defs, "-c", s, "-o", o])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
        return r

    if todo:
        with ThreadPoolExecutor(max_workers=jobs or os.cpu_count() or 4) as ex:
            list(ex.map(run, todo))
    stale = not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if todo or stale or force:
        run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
