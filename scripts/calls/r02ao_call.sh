#!/bin/bash
# 1 GPU: ncu --set full of the two single-purpose persistent kernels (each after the same command exited 0 without ncu)
OUT=gpurun_out
mkdir -p $OUT
CMD1="python bench.py --workload cfg1 --dtype f32 --steps 2000 --warmup 10 --no-cpu-baseline"
$CMD1 > $OUT/plain_small_r02f.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_train_small -s 1 -c 1 -o $OUT/prof_k_train_small_r02f -f $CMD1 > $OUT/ncu_k_train_small_r02f.log 2>&1
tail -2 $OUT/ncu_k_train_small_r02f.log
CMD2="python scripts/sample_bench.py 1024 20000"
$CMD2 > $OUT/plain_k9_r02f.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_recur_persist -s 1 -c 1 -o $OUT/prof_k_recur_persist_r02f -f $CMD2 > $OUT/ncu_k_recur_persist_r02f.log 2>&1
tail -2 $OUT/ncu_k_recur_persist_r02f.log
ls -la $OUT | grep -E "prof_k_(train_small|recur_persist)_r02f"
