#!/bin/bash
# 1 GPU, the build that is committed last: smoke, the whole GPU suite, the default bench and the reference arm
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee $OUT/r02aq_smoke.txt
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $OUT/r02aq_pytest_gpu.txt
timeout 900 python bench.py > $OUT/r02aq_bench_default.json 2> $OUT/r02aq_bench_default.err; tail -2 $OUT/r02aq_bench_default.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02aq_bench_20steps.json 2> $OUT/r02aq_bench_20steps.err
python - <<'PY'
import json
for f in ("default", "20steps"):
    d = json.loads(open(f"gpurun_out/r02aq_bench_{f}.json").read().strip().splitlines()[-1])
    print(f, d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "launches", d["gpu_launches"], "loss", d["final_loss_bits_per_char"], d["us_per_recurrent_timestep"], d["clocks"])
PY
