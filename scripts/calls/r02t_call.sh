#!/bin/bash
# 2 GPUs: tuning matrix for the data-parallel overlap (NCCL CTA cap x K6a panel layout), one bench line each
OUT=gpurun_out
mkdir -p $OUT
i=0
for V in "16 442" "16 64" "8 442" "32 442" "64 64" "16 82" "24 442"; do
  set -- $V
  i=$((i+1))
  LSTM_TUNE_NCCL_CTAS=$1 LSTM_TUNE_PANELS=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29540+i)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02t_$1_$2.json 2> $OUT/r02t_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
f = f"gpurun_out/r02t_{sys.argv[1]}_{sys.argv[2]}.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    p = d["phases_ms_last_step"]
    print(sys.argv[1:], "ms", round(d["ms_per_step"], 3), "wgrad", round(p["weight_grads"], 3), "wait", round(p["allreduce_wait"], 3), "total", round(p["total"], 3), "ok", d["dp_check"]["ok"])
    print("   ", p.get("comm_timeline"))
except Exception as e:
    print(sys.argv[1:], "FAILED", e)
PY
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02t_1gpu.json 2> $OUT/r02t_1gpu.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02t_1gpu.json").read().strip().splitlines()[-1]); print("1gpu ms", d["ms_per_step"], d["phases_ms_last_step"])
PY
