#!/bin/bash
# 1 GPU: the one-kernel training path for the reference's default shape: bit-identity tests, the config-1 oracle tests, bench
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_train_small.py -x -q -m gpu 2>&1 | tail -15 | tee $OUT/r02u_pytest_small.txt
timeout 1200 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_lstm_binary.py tests/test_gpu_options.py -q -m gpu 2>&1 | tail -8 | tee $OUT/r02u_pytest_f32.txt
timeout 300 python bench.py --workload cfg1 --dtype f32 --steps 20000 --warmup 200 > $OUT/r02u_bench_cfg1.json 2> $OUT/r02u_bench_cfg1.err; tail -3 $OUT/r02u_bench_cfg1.err; cat $OUT/r02u_bench_cfg1.json | cut -c1-1500
