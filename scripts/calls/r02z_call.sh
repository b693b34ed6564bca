#!/bin/bash
# 2 GPUs: data-parallel iteration as graph SEGMENTS (default) vs plain stream launches (LSTM_NO_GRAPH=1), same box, twice each
OUT=gpurun_out
mkdir -p $OUT
i=0
for MODE in graph plain graph plain; do
  i=$((i+1))
  if [ $MODE = plain ]; then export LSTM_NO_GRAPH=1; else unset LSTM_NO_GRAPH; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29560+i)) bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02z_${MODE}_$i.json 2> $OUT/r02z_${MODE}_$i.err; tail -2 $OUT/r02z_${MODE}_$i.err
  python - $MODE $i <<'PY'
import json, sys
f = f"gpurun_out/r02z_{sys.argv[1]}_{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "profile total", round(d["phases_ms_last_step"]["total"], 3), "launches", d["gpu_launches"], "ok", d["dp_check"]["ok"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
