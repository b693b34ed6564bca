#!/bin/bash
# 2 GPUs: data-parallel correctness tests + the bench with its dp_check, single-GPU bench for reference
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_dp.py -q 2>&1 | tail -8 | tee $OUT/r02j_pytest_dp.txt
timeout 300 python bench.py --steps 10 --warmup 3 > $OUT/r02j_bench_1gpu.json 2> $OUT/r02j_bench_1gpu.err; tail -3 $OUT/r02j_bench_1gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/r02j_bench_2gpu.json 2> $OUT/r02j_bench_2gpu.err; tail -5 $OUT/r02j_bench_2gpu.err
python - <<'PY'
import json
for f in ("gpurun_out/r02j_bench_1gpu.json", "gpurun_out/r02j_bench_2gpu.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["e2e"]); print(d.get("f32")); print(d.get("cpu_baseline")); print(d.get("dp_check")); print(d["phases_ms_last_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
