#!/bin/bash
# 2 GPUs: DP bench with the 8+2 panel split and the moved small bucket; then the DP tests; then 1-GPU bench of the same build
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02r_bench_2gpu.json 2> $OUT/r02r_bench_2gpu.err; tail -3 $OUT/r02r_bench_2gpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02r_bench_1gpu.json 2> $OUT/r02r_bench_1gpu.err; tail -3 $OUT/r02r_bench_1gpu.err
timeout 900 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -3 | tee $OUT/r02r_pytest_dp.txt
python - <<'PY'
import json
for f in ("r02r_bench_2gpu", "r02r_bench_1gpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d.get("dp_check")); print(d["phases_ms_last_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
