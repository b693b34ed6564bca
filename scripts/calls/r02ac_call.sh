#!/bin/bash
# 1 GPU: K9 with 128-bit shared-memory loads: sampling / evaluation tests, cfg5 bench twice
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_edge_cases.py tests/test_gpu_bench_variants.py -q -m gpu -k "persistent or sampl or eval or bpc or beside or enwik5 or untrained" 2>&1 | tail -4 | tee $OUT/r02ac_pytest.txt
for i in 1 2; do
  timeout 600 python bench.py --workload cfg5 > $OUT/r02ac_bench_cfg5_$i.json 2> $OUT/r02ac_bench_cfg5_$i.err; tail -2 $OUT/r02ac_bench_cfg5_$i.err
  python - $i <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02ac_bench_cfg5_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print("cfg5", d["value"], d["us_per_sampled_char"], d["us_per_evaluated_char"], d["eval_bits_per_char"])
PY
done
