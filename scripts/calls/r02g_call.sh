#!/bin/bash
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r02g_pytest_all.txt
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r02g_bench.json 2> $OUT/r02g_bench.err; cut -c1-400 $OUT/r02g_bench.json
