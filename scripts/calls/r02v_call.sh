#!/bin/bash
# 1 GPU: small-kernel tests + clock stamps; then 2-GPU-free checks of the new window / one-hot kernels via the bench-variant tests
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train_small.py -q -m gpu 2>&1 | tail -8 | tee $OUT/r02v_pytest_small.txt
python scripts/small_clocks.py 64 3 2>&1 | tee $OUT/r02v_small_clocks.txt
python scripts/small_clocks.py 64 5 2>&1 | tee -a $OUT/r02v_small_clocks.txt
timeout 1500 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_lstm_binary.py -q -m gpu -x 2>&1 | tail -5 | tee $OUT/r02v_pytest_variants.txt
timeout 300 python bench.py --workload cfg1 --dtype f32 --steps 20000 --warmup 200 > $OUT/r02v_bench_cfg1.json 2> $OUT/r02v_bench_cfg1.err; tail -2 $OUT/r02v_bench_cfg1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02v_bench_cfg1.json").read().strip().splitlines()[-1]); print("1gpu", d["value"], d["ms_per_step"], d["e2e"]["value"], d["phases_ms_last_step"])
PY
