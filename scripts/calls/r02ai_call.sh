#!/bin/bash
# 1 GPU, final build: smoke, full GPU test-suite, the default bench line (with f32 sub-object and CPU baseline), the reference arm,
# and the other workloads
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/r02ai_smoke.txt
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee $OUT/r02ai_pytest_gpu.txt
timeout 900 python bench.py > $OUT/r02ai_bench_default.json 2> $OUT/r02ai_bench_default.err; tail -2 $OUT/r02ai_bench_default.err
timeout 900 python bench.py --impl reference > $OUT/r02ai_bench_reference.json 2> $OUT/r02ai_bench_reference.err; tail -2 $OUT/r02ai_bench_reference.err
for WL in cfg3 cfg2; do
  timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02ai_bench_$WL.json 2> $OUT/r02ai_bench_$WL.err; tail -2 $OUT/r02ai_bench_$WL.err
done
timeout 300 python bench.py --workload cfg1 --dtype f32 --steps 20000 --warmup 200 > $OUT/r02ai_bench_cfg1.json 2> $OUT/r02ai_bench_cfg1.err; tail -2 $OUT/r02ai_bench_cfg1.err
timeout 600 python bench.py --workload cfg5 > $OUT/r02ai_bench_cfg5.json 2> $OUT/r02ai_bench_cfg5.err; tail -2 $OUT/r02ai_bench_cfg5.err
python - <<'PY'
import json
for f in ("default", "reference", "cfg3", "cfg2", "cfg1", "cfg5"):
    try:
        d = json.loads(open(f"gpurun_out/r02ai_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("unit"), "ms", d.get("ms_per_step"), "e2e", d.get("e2e", {}).get("value"), "roofline", {k: d.get("roofline", {}).get(k) for k in ("kernel", "achieved", "frac")},
              "cpu", d.get("cpu_baseline", {}).get("value"), "f32", (d.get("f32") or {}).get("value"), d.get("us_per_recurrent_timestep"), d.get("us_per_sampled_char"), d.get("us_per_evaluated_char"))
    except Exception as e:
        print(f, "FAILED", e)
PY
