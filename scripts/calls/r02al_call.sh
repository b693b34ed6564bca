#!/bin/bash
# 1 GPU: vectorised operand refresh: bf16 tests, bench lines (the Adagrad phase), a learning run of the drop-in binary on enwik6
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_bf16.py tests/test_gpu_options.py -x -q -m gpu 2>&1 | tail -4 | tee $OUT/r02al_pytest.txt
for WL in cfg4 cfg3 cfg2; do
  timeout 600 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu-baseline --no-f32 > $OUT/r02al_bench_$WL.json 2> $OUT/r02al_bench_$WL.err; tail -2 $OUT/r02al_bench_$WL.err
  python - $WL <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/r02al_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], "adagrad phase", d["phases_ms_last_step"]["adagrad"], "loss", d["final_loss_bits_per_char"])
PY
done
cd eigen_lstm_b200 && timeout 600 ./lstm --file ../tests/golden/enwik6.txt --hidden 512 --seq 100 --batch 64 --stride 99 --forget-bias 1 --bf16 --epochs 3 --lr 0.01 --train-percent 95 --test-every 2 --state-std 0 --seed 1 2>&1 | grep -v "%" | tail -25 | tee ../$OUT/r02al_cfg2_enwik6_run.txt
