#!/bin/bash
# 1 GPU: the last 3/16 of the weight-gradient GEMM's K range beside BPTT: tests, then A/B on one box
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bench_variants.py -x -q -m gpu -k "beside or agree or training" 2>&1 | tail -3 | tee $OUT/r02af_pytest.txt
i=0
for MODE in split nosplit split nosplit; do
  i=$((i+1))
  if [ $MODE = nosplit ]; then export LSTM_TUNE_NO_K6SPLIT=1; else unset LSTM_TUNE_NO_K6SPLIT; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02af_${MODE}_$i.json 2> $OUT/r02af_${MODE}_$i.err; tail -2 $OUT/r02af_${MODE}_$i.err
  python - $MODE $i <<'PY'
import json, sys
f = f"gpurun_out/r02af_{sys.argv[1]}_{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d["phases_ms_last_step"]
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "profile total", round(p["total"], 3), "bwd", round(p["bwd_recurrence"], 3), "wgrad", round(p["weight_grads"], 3), "clk", d["clocks"]["sm_mhz"], "loss", d["final_loss_bits_per_char"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
