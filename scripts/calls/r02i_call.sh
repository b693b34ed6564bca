#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/r02i_pytest_all.txt
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r02i_bench.json 2> $OUT/r02i_bench.err; tail -3 $OUT/r02i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02i_bench.json"))
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["e2e"]); print(d.get("f32")); print(d.get("cpu_baseline")); print(d["roofline"])
PY
timeout 200 python bench.py --workload cfg5 --steps 2 > $OUT/r02i_bench_cfg5.json 2> $OUT/r02i_bench_cfg5.err; cut -c1-900 $OUT/r02i_bench_cfg5.json; tail -3 $OUT/r02i_bench_cfg5.err
