#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py -q -x 2>&1 | tail -4 | tee $OUT/r02q_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02q_bench.json 2> $OUT/r02q_bench.err; tail -2 $OUT/r02q_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02q_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["phases_ms_last_step"]); print(d["roofline"])
PY
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -14 | tee $OUT/r02q_clocks_cfg4.txt
for W in cfg2 cfg3; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02q_bench_$W.json 2> $OUT/r02q_bench_$W.err
  python - $W <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02q_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"])
PY
done
