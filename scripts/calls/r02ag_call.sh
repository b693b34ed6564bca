#!/bin/bash
# 1 GPU: which share of the weight-gradient GEMM's K range (in sixteenths of the window) should run beside BPTT?
OUT=gpurun_out
mkdir -p $OUT
for L in 0 2 3 4 5 0 2 3 4 5; do
  export LSTM_TUNE_K6_LATE=$L
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ag_$L.json 2> $OUT/r02ag_$L.err; tail -2 $OUT/r02ag_$L.err
  python - $L <<'PY'
import json, sys
f = f"gpurun_out/r02ag_{sys.argv[1]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d["phases_ms_last_step"]
    print("late", sys.argv[1], "/16  ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "profile wgrad", round(p["weight_grads"], 3), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
