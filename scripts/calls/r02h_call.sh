#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/r02h_pytest_all.txt
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02h_bench.json 2> $OUT/r02h_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02h_bench.json"))
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["phases_ms_last_step"]); print(d["e2e"])
PY
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -26 | tee $OUT/r02h_clocks.txt
