#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee $OUT/r02l_pytest_all.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02l_bench.json 2> $OUT/r02l_bench.err; tail -3 $OUT/r02l_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["phases_ms_last_step"])
PY
for W in cfg2 cfg3; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02l_bench_$W.json 2> $OUT/r02l_bench_$W.err
  python - $W <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02l_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"], d["phases_ms_last_step"])
PY
done
BPC_OUT=r02l_bpc_cfg3shape_enwik6.json timeout 900 python scripts/bpc_bf16_vs_f32.py 1024 128 101 3000 0.01 2>&1 | tail -5 | tee $OUT/r02l_bpc.txt
