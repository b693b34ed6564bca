#!/bin/bash
# 8 GPUs of one box: the data-parallel bench with its dp_check (every rank), max-over-ranks timing
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/r02o_bench_8gpu.json 2> $OUT/r02o_bench_8gpu.err; tail -4 $OUT/r02o_bench_8gpu.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02o_bench_8gpu.json").read().strip().splitlines()[-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(d.get("dp_check")); print(d["phases_ms_last_step"])
except Exception as e:
    print("FAILED", e)
PY
