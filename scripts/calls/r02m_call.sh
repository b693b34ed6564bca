#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_parity_f32.py tests/test_gpu_edge_cases.py -q 2>&1 | tail -6 | tee $OUT/r02m_pytest.txt
timeout 200 python bench.py --workload cfg5 --steps 1 > $OUT/r02m_bench_cfg5.json 2> $OUT/r02m_bench_cfg5.err; cut -c1-900 $OUT/r02m_bench_cfg5.json | tr ',' '\n' | grep "us_per\|value"; tail -2 $OUT/r02m_bench_cfg5.err
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg3 2>&1 | tail -26 | tee $OUT/r02m_clocks_cfg3.txt
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -26 | tee $OUT/r02m_clocks_cfg4.txt
( cd $OUT && timeout 300 ../eigen_lstm_b200/lstm --file ../tests/golden/enwik6.txt --hidden 512 --seq 100 --batch 64 --stride 99 --forget-bias 1 --bf16 \
    --epochs 12 --train-percent 95 --test-every 2 --sample 300 --lr 0.01 --state-std 0 --seed 1 --progress eta --save r02m_cfg2run 2>&1 | tr '\r' '\n' | grep -v "^ *\[Epoch" | tail -60 > r02m_cfg2_enwik6_run.txt; tail -25 r02m_cfg2_enwik6_run.txt | cut -c1-200 )
BPC_OUT=r02m_bpc_cfg3shape_enwik6_lr0.002.json timeout 900 python scripts/bpc_bf16_vs_f32.py 1024 128 101 3000 0.002 2>&1 | tail -5 | tee $OUT/r02m_bpc.txt
