#!/bin/bash
# 1 GPU: repeatability of the final build: the GPU test-suite three times, the bench three times (same loss bits every time)
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
  timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -2
done | tee $OUT/r02am_pytest_x3.txt
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02am_bench_$i.json 2> $OUT/r02am_bench_$i.err
  python - $i <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02am_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print("run", sys.argv[1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "loss", repr(d["final_loss_bits_per_char"]), d["clocks"])
PY
done | tee $OUT/r02am_bench_x3.txt
