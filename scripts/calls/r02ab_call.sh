#!/bin/bash
# 1 GPU: A/B on one box: K3 / K6c beside the recurrences (programmatic dependents) vs one after the other
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_bench_variants.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/r02ab_pytest.txt
i=0
for MODE in beside serial beside serial beside serial; do
  i=$((i+1))
  if [ $MODE = serial ]; then export LSTM_TUNE_NO_OVERLAP=1; else unset LSTM_TUNE_NO_OVERLAP; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ab_${MODE}_$i.json 2> $OUT/r02ab_${MODE}_$i.err; tail -2 $OUT/r02ab_${MODE}_$i.err
  python - $MODE $i <<'PY'
import json, sys
f = f"gpurun_out/r02ab_{sys.argv[1]}_{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "profile total", round(d["phases_ms_last_step"]["total"], 3), "clk", d["clocks"]["sm_mhz"], "loss", d["final_loss_bits_per_char"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
unset LSTM_TUNE_NO_OVERLAP
for WL in cfg3 cfg2; do
  timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ab_bench_$WL.json 2> $OUT/r02ab_bench_$WL.err
  python - $WL <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02ab_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"])
PY
done
