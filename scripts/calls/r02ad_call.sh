#!/bin/bash
# 8 GPUs: data-parallel bench, K6a in 6+4 panels (default) vs 4+4+2 panels
OUT=gpurun_out
mkdir -p $OUT
i=0
for MODE in p64 p442; do
  i=$((i+1))
  if [ $MODE = p442 ]; then export LSTM_TUNE_PANELS3=1; else unset LSTM_TUNE_PANELS3; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29570+i)) bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ad_${MODE}_8gpu.json 2> $OUT/r02ad_${MODE}_8gpu.err; tail -2 $OUT/r02ad_${MODE}_8gpu.err
  python - $MODE <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/r02ad_{sys.argv[1]}_8gpu.json").read().strip().splitlines()[-1])
    p = d["phases_ms_last_step"]
    print(sys.argv[1], d["value"], "ms", round(d["ms_per_step"], 3), "e2e", d["e2e"]["value"], "wgrad", round(p["weight_grads"], 3), "wait", round(p["allreduce_wait"], 3), "total", round(p["total"], 3), d["dp_check"])
    print("   ", p.get("comm_timeline"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
