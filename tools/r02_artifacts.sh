#!/bin/bash
# Round-2 evidence run on one B200: bench lines (both arms), ncu launch list of the bench command, ncu --set full of the
# two HBM-bound round-2 kernels.  Outputs in gpurun_out/r02_*; the summaries are copied to profiles/ afterwards.
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py > $OUT/r02_bench_full.json 2> $OUT/r02_bench_full.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference.json 2> $OUT/r02_bench_reference.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/r02_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_onepass -s 2 -c 1 -f -o $OUT/r02_ncu_k3_onepass \
    python tools/bench_scan.py 26 > $OUT/r02_ncu_k3_onepass.log 2>&1; echo "ncu k3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1b_fifo_jobs -s 1 -c 1 -f -o $OUT/r02_ncu_k1b \
    python tools/bench_k1b.py > $OUT/r02_ncu_k1b.log 2>&1; echo "ncu k1b rc=$?"
tail -c 600 $OUT/r02_bench_full.json
