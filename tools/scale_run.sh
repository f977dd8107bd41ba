#!/bin/bash
# Scaling runs of every BASELINE configuration on the GPUs of one box (driver launch contract: torchrun, one rank per GPU).
#   tools/scale_run.sh "1 2 4 8"   -> gpurun_out/r02_scale_<config>_<scaling>.jsonl (one JSON line per N)
NS=${1:-"1 2 4 8"}
OUT=gpurun_out
mkdir -p $OUT
PORT=29600
run() {  # config scaling N extra...
  local cfg=$1 sc=$2 n=$3; shift 3
  PORT=$((PORT+1))
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --config $cfg --scaling $sc --no-cpu --no-others --no-replay --no-fp64 "$@" 2>>$OUT/r02_scale.err | tail -1 >> $OUT/r02_scale_${cfg}_${sc}.jsonl
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n --config $cfg --scaling $sc --no-cpu --no-others --no-replay --no-fp64 "$@" 2>>$OUT/r02_scale.err | grep '^{' | tail -1 >> $OUT/r02_scale_${cfg}_${sc}.jsonl
  fi
}
rm -f $OUT/r02_scale_*.jsonl $OUT/r02_scale.err
for n in $NS; do
  run cfg2 strong $n --steps 5 --warmup 3
  run cfg3 strong $n --steps 3 --warmup 3
  run cfg5 strong $n --steps 3 --warmup 3
done
run cfg5 weak $(echo $NS | awk '{print $NF}') --steps 3 --warmup 3
run cfg2 weak $(echo $NS | awk '{print $NF}') --steps 3 --warmup 3
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $OUT/r02_mgpu_pytest.log 2>&1
tail -3 $OUT/r02_mgpu_pytest.log
for f in $OUT/r02_scale_*.jsonl; do echo $f; python - "$f" <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    d = json.loads(ln)
    print("  N=%d value=%.4g e2e=%.4g ms/step=%.1f scaling=%s frac=%.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["scaling"], d["roofline"]["frac"]))
PY
done
