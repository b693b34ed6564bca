#!/bin/bash
# 1 GPU: K3 beside the forward recurrence, K6c beside BPTT (side stream): tests, then bench cfg4 / cfg3 / cfg2
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py tests/test_gpu_options.py tests/test_gpu_lstm_binary.py -x -q -m gpu 2>&1 | tail -8 | tee $OUT/r02aa_pytest.txt
for WL in cfg4 cfg3 cfg2; do
  timeout 600 python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02aa_bench_$WL.json 2> $OUT/r02aa_bench_$WL.err; tail -2 $OUT/r02aa_bench_$WL.err
done
python - <<'PY'
import json
for f in ("cfg4", "cfg3", "cfg2"):
    try:
        d = json.loads(open(f"gpurun_out/r02aa_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["final_loss_bits_per_char"], d["phases_ms_last_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
