#!/bin/bash
# scripts/profile_gpu.sh — run under gpurun on ONE B200.  Produces in gpurun_out/:
#   launches_<tag>.csv             every launch of the bench command with its device time
#                                  (ncu, cold-cache, serialised: compare SHARES, not absolutes)
#   prof_<kernel>_<tag>.ncu-rep    one --set full capture per hot kernel
# Each ncu pass runs only after the same command exited 0 without ncu (B200_PROFILING.md).
# LSTM_NO_GRAPH=1: the same kernels as plain stream launches (ncu then sees ordinary launches).
TAG=${1:-r02}
WL=${2:-cfg4}
export LSTM_NO_GRAPH=1
CMD="python bench.py --workload $WL --dtype bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-f32"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-400} --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in ${KERNELS:-k_bwd_recur k_fwd_recur k_gemm_nt_pair k_logits k_adagrad_f32}; do
  case $K in k_gemm_nt*) KS=7;; *) KS=3;; esac
  $CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $KS -c 1 \
      -o gpurun_out/prof_${K}_$TAG -f $CMD > gpurun_out/ncu_${K}_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_${K}_$TAG.log
done
ls -la gpurun_out/ | grep $TAG
