#!/bin/bash
# 1 GPU: contention experiment — BPTT timestep anatomy (timesteps 40 / 80 / 120 of 256) with and without 20 % of K6a running beside it
OUT=gpurun_out
mkdir -p $OUT
export LSTM_TC_DEBUG=1
for STEP in 40 80 120; do
  export LSTM_TC_DEBUG_STEP=$STEP
  for MODE in alone k6a; do
    if [ $MODE = k6a ]; then export LSTM_TUNE_K6A_BESIDE=1; else unset LSTM_TUNE_K6A_BESIDE; fi
    echo "== step $STEP $MODE"; python scripts/recur_clocks.py cfg4 2>&1 | grep -A 12 "BPTT recurrence" | grep -E "period|detail"
  done
done | tee $OUT/r02ae_k6a_beside_bptt_clocks2.txt
