#!/bin/bash
# 1 GPU: one-kernel path through lstm_train_step with the window as a kernel argument and the loss written to the pinned ring
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_train_small.py tests/test_gpu_parity_f32.py tests/test_gpu_lstm_binary.py tests/test_gpu_edge_cases.py -q -m gpu 2>&1 | tail -5 | tee $OUT/r02aj_pytest.txt
timeout 300 python bench.py --workload cfg1 --dtype f32 --steps 20000 --warmup 200 > $OUT/r02aj_bench_cfg1.json 2> $OUT/r02aj_bench_cfg1.err; tail -2 $OUT/r02aj_bench_cfg1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02aj_bench_cfg1.json").read().strip().splitlines()[-1]); print("cfg1", d["value"], d["ms_per_step"], d["e2e"])
PY
