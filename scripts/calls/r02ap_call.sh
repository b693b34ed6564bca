#!/bin/bash
# 1 GPU: one-hot build beside the forward recurrence: bf16 tests + bench
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py tests/test_gpu_options.py tests/test_gpu_lstm_binary.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/r02ap_pytest.txt
for i in 1 2; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ap_bench_$i.json 2> $OUT/r02ap_bench_$i.err
  python - $i <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02ap_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print("run", sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", repr(d["final_loss_bits_per_char"]), d["gpu_launches"])
PY
done
