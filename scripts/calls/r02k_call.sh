#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py -x -q 2>&1 | tail -5 | tee $OUT/r02k_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02k_bench.json 2> $OUT/r02k_bench.err; tail -3 $OUT/r02k_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["phases_ms_last_step"])
PY
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -26 | tee $OUT/r02k_clocks.txt
timeout 200 python bench.py --workload cfg5 --steps 1 > $OUT/r02k_bench_cfg5.json 2> $OUT/r02k_bench_cfg5.err; cut -c1-700 $OUT/r02k_bench_cfg5.json | tr ',' '\n' | grep "us_per"
