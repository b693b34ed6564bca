#!/bin/bash
# 1 GPU: full GPU test-suite, small-kernel clocks (typical iteration), bench lines cfg1 / cfg5 / default
OUT=gpurun_out
mkdir -p $OUT
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee $OUT/r02w_pytest_gpu.txt
python scripts/small_clocks.py 64 3 2>&1 | tee $OUT/r02w_small_clocks.txt
python scripts/small_clocks.py 64 5 2>&1 | tee -a $OUT/r02w_small_clocks.txt
timeout 300 python bench.py --workload cfg1 --dtype f32 --steps 20000 --warmup 200 > $OUT/r02w_bench_cfg1.json 2> $OUT/r02w_bench_cfg1.err; tail -2 $OUT/r02w_bench_cfg1.err
timeout 600 python bench.py --workload cfg5 > $OUT/r02w_bench_cfg5.json 2> $OUT/r02w_bench_cfg5.err; tail -2 $OUT/r02w_bench_cfg5.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02w_bench_cfg4.json 2> $OUT/r02w_bench_cfg4.err; tail -2 $OUT/r02w_bench_cfg4.err
python - <<'PY'
import json
for f in ("cfg1", "cfg5", "cfg4"):
    try:
        d = json.loads(open(f"gpurun_out/r02w_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["unit"], d["ms_per_step"], d.get("e2e", {}).get("value"), d.get("phases_ms_last_step"), d.get("sampling"), d.get("evaluation"))
    except Exception as e:
        print(f, "FAILED", e)
PY
