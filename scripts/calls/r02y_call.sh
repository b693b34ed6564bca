#!/bin/bash
# 2 GPUs: is the 2-GPU number bounded by the slower GPU of the box?  1-GPU bench on each GPU, then the DP bench
OUT=gpurun_out
mkdir -p $OUT
for G in 0 1; do
  CUDA_VISIBLE_DEVICES=$G timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02y_bench_gpu$G.json 2> $OUT/r02y_bench_gpu$G.err; tail -2 $OUT/r02y_bench_gpu$G.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02y_bench_2gpu.json 2> $OUT/r02y_bench_2gpu.err; tail -3 $OUT/r02y_bench_2gpu.err
python - <<'PY'
import json
for f in ("r02y_bench_gpu0", "r02y_bench_gpu1", "r02y_bench_2gpu"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]); print(d["phases_ms_last_step"])
    except Exception as e:
        print(f, "FAILED", e)
PY
