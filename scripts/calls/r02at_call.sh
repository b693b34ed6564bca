#!/bin/bash
# N GPUs (N = $1), final build: DP bench line (20 steps), the DP tests (N = 2 only)
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02at_bench_${N}gpu.json 2> $OUT/r02at_bench_${N}gpu.err; tail -3 $OUT/r02at_bench_${N}gpu.err
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -3 | tee $OUT/r02at_pytest_dp.txt; fi
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r02at_bench_{n}gpu.json").read().strip().splitlines()[-1])
    print(n, "GPUs", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"]); print(d.get("dp_check")); print(d["phases_ms_last_step"])
except Exception as e:
    print(n, "FAILED", e)
PY
