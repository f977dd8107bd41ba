#!/bin/bash
# cfg3 (one Gamma/Pareto/1 chain of 1e10 jobs cut into N ranges) on 1, 2, 4, 8 GPUs of one box + the multi-GPU tests
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/r02_scale_cfg3_strong_v2.jsonl
PORT=29700
for n in 1 2 4 8; do
  PORT=$((PORT+1))
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --config cfg3 --scaling strong --no-cpu --no-others --no-replay --no-fp64 --steps 3 --warmup 3 2>>$OUT/r02_scale3.err | tail -1 >> $OUT/r02_scale_cfg3_strong_v2.jsonl
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n --config cfg3 --scaling strong --no-cpu --no-others --no-replay --no-fp64 --steps 3 --warmup 3 2>>$OUT/r02_scale3.err | grep '^{' | tail -1 >> $OUT/r02_scale_cfg3_strong_v2.jsonl
  fi
done
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/r02_mgpu_pytest_v2.log 2>&1
tail -2 $OUT/r02_mgpu_pytest_v2.log
python - <<'PY'
import json
for ln in open("gpurun_out/r02_scale_cfg3_strong_v2.jsonl"):
    d = json.loads(ln)
    print("N=%d value=%.4g ms/step=%.1f" % (d["n_gpus"], d["value"], d["ms_per_step"]))
PY
