#!/bin/bash
# ncu evidence for profiles/: launch list + --set full of the hot kernels (one ncu family per call), after plain runs exited 0
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py -q -x 2>&1 | tail -4 | tee $OUT/r02n_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02n_bench.json 2> $OUT/r02n_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02n_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["us_per_recurrent_timestep"]); print(d["phases_ms_last_step"])
PY
bash scripts/profile_gpu.sh r02 cfg4 2>&1 | tail -20
BPC_MEM0=1 BPC_OUT=r02n_bpc_cfg3shape_enwik6_mem0.json timeout 900 python scripts/bpc_bf16_vs_f32.py 1024 128 101 3000 0.01 2>&1 | tail -5 | tee $OUT/r02n_bpc.txt
