#!/bin/bash
# compute-sanitizer memcheck over a small run that touches every kernel family (one tool per call, after a plain run exited 0)
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/sanitizer_run.py > $OUT/r02p_plain.txt 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitizer_run.py > $OUT/r02p_memcheck.txt 2>&1
echo "memcheck exit $?" | tee -a $OUT/r02p_memcheck.txt
tail -15 $OUT/r02p_memcheck.txt
# K9 after the bank-conflict / split-barrier changes + clock stamps with the finer BPTT detail
timeout 200 python bench.py --workload cfg5 --steps 1 > $OUT/r02p_bench_cfg5.json 2> $OUT/r02p_bench_cfg5.err; cut -c1-900 $OUT/r02p_bench_cfg5.json | tr ',' '\n' | grep "us_per\|\"value"; tail -2 $OUT/r02p_bench_cfg5.err
timeout 600 python -m pytest tests/test_gpu_parity_f32.py tests/test_gpu_edge_cases.py tests/test_gpu_lstm_binary.py -q 2>&1 | tail -4
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -27 | tee $OUT/r02p_clocks_cfg4.txt
