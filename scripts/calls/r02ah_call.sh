#!/bin/bash
# 1 GPU: weight-gradient GEMM as cta_group::2 pairs (256 x 256 tiles) vs single-CTA 128 x 256 tiles: tests, then A/B on one box
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py -x -q -m gpu 2>&1 | tail -5 | tee $OUT/r02ah_pytest.txt
i=0
for MODE in pair single pair single; do
  i=$((i+1))
  if [ $MODE = single ]; then export LSTM_TUNE_NO_K6PAIR=1; else unset LSTM_TUNE_NO_K6PAIR; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ah_${MODE}_$i.json 2> $OUT/r02ah_${MODE}_$i.err; tail -2 $OUT/r02ah_${MODE}_$i.err
  python - $MODE $i <<'PY'
import json, sys
f = f"gpurun_out/r02ah_{sys.argv[1]}_{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d["phases_ms_last_step"]
    print(sys.argv[1], "ms", round(d["ms_per_step"], 3), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "profile total", round(p["total"], 3), "wgrad", round(p["weight_grads"], 3), "clk", d["clocks"]["sm_mhz"], "loss", d["final_loss_bits_per_char"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
