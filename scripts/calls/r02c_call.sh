#!/bin/bash
# round 2, call c: new parity / option tests, then steady-state DRAM traffic of the step kernels with ncu in APPLICATION
# replay mode (no memory save/restore, no cache flush: what the kernels see inside a real iteration)
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r02c_pytest.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
export LSTM_NO_GRAPH=1
$CMD > $OUT/r02c_plain.log 2>&1 &&
ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_fwd_step -s 900 -c 4 --csv --log-file $OUT/r02c_ncu_fwd_appreplay.csv $CMD > $OUT/r02c_ncu_fwd.log 2>&1
ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_bwd_step -s 900 -c 4 --csv --log-file $OUT/r02c_ncu_bwd_appreplay.csv $CMD > $OUT/r02c_ncu_bwd.log 2>&1
LSTM_L2HINT=1 ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_bwd_step -s 900 -c 4 --csv --log-file $OUT/r02c_ncu_bwd_appreplay_hint1.csv $CMD > $OUT/r02c_ncu_bwd_h1.log 2>&1
grep -h "dram__bytes\|duration\|hit_rate" $OUT/r02c_ncu_*appreplay*.csv | cut -d, -f5,13- | cut -c1-200
