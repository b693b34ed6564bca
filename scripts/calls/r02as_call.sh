#!/bin/bash
# 1 GPU: bench.py --workload cfg5 after the evaluation timing change (median of three)
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python bench.py --workload cfg5 --steps 2 > $OUT/r02as_bench_cfg5.json 2> $OUT/r02as_bench_cfg5.err; tail -2 $OUT/r02as_bench_cfg5.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02as_bench_cfg5.json").read().strip().splitlines()[-1]); print("cfg5", d["value"], d["us_per_sampled_char"], d["us_per_evaluated_char"], d["us_per_evaluated_char_runs"])
PY
