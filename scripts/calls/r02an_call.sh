#!/bin/bash
# 1 GPU: bf16 vs fp32 bits-per-char on the full enwik6 at config 3's shape with the FINAL build (Adagrad initial accumulator 1)
OUT=gpurun_out
mkdir -p $OUT
BPC_MEM0=1 BPC_OUT=$OUT/r02an_bpc_cfg3shape_enwik6_mem0_final.json timeout 1500 python scripts/bpc_bf16_vs_f32.py 1024 128 101 3000 0.01 2>&1 | tail -12 | tee $OUT/r02an_bpc.txt
