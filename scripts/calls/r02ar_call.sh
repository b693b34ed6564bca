#!/bin/bash
# 1 GPU, committed build: the other workloads' bench lines and the reference arm
OUT=gpurun_out
mkdir -p $OUT
for WL in cfg3 cfg2; do
  timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ar_bench_$WL.json 2> $OUT/r02ar_bench_$WL.err; tail -1 $OUT/r02ar_bench_$WL.err
done
timeout 600 python bench.py --workload cfg5 > $OUT/r02ar_bench_cfg5.json 2> $OUT/r02ar_bench_cfg5.err; tail -1 $OUT/r02ar_bench_cfg5.err
timeout 900 python bench.py --impl reference > $OUT/r02ar_bench_reference.json 2> $OUT/r02ar_bench_reference.err; tail -1 $OUT/r02ar_bench_reference.err
python - <<'PY'
import json
for f in ("cfg3", "cfg2", "cfg5", "reference"):
    try:
        d = json.loads(open(f"gpurun_out/r02ar_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), "ms", d.get("ms_per_step"), d.get("us_per_recurrent_timestep"), d.get("us_per_sampled_char"), d.get("us_per_evaluated_char"), (d.get("cpu_baseline") or {}).get("cores"))
    except Exception as e:
        print(f, "FAILED", e)
PY
