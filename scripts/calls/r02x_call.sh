#!/bin/bash
# N GPUs (N = $1): DP bench (dp_check, comm timeline), the DP tests, and a 1-GPU bench of the same build on the same box
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02x_bench_${N}gpu.json 2> $OUT/r02x_bench_${N}gpu.err; tail -3 $OUT/r02x_bench_${N}gpu.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02x_bench_1gpu_box${N}.json 2> $OUT/r02x_bench_1gpu_box${N}.err; tail -3 $OUT/r02x_bench_1gpu_box${N}.err
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -3 | tee $OUT/r02x_pytest_dp.txt; fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for f in (f"r02x_bench_{n}gpu", f"r02x_bench_1gpu_box{n}"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d.get("dp_check")); print(d["phases_ms_last_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
