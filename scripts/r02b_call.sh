#!/bin/bash
# round 2, call b: (1) L2 eviction-priority hints on the step kernels' TMA loads, (2) steady-state DRAM traffic of the step
# kernels (ncu WITHOUT its default cache flush between kernels: --cache-control none)
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
line() {
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    u = d["us_per_recurrent_timestep"]
    print(f'{d["value"]:12.0f} chars/s {d["ms_per_step"]:8.3f} ms/step  fwd {u["forward"]:6.2f} us  bwd {u["backward"]:6.2f} us  '
          f'loss {d["final_loss_bits_per_char"]!r}  launches {d["gpu_launches"]}')
except Exception as e:
    print("FAILED:", e)
PY
}
for V in "LSTM_L2HINT=0" "LSTM_L2HINT=1" "LSTM_L2HINT=2" "LSTM_L2HINT=3" "LSTM_L2HINT=1 LSTM_NO_L2PIN=1" "LSTM_L2HINT=3 LSTM_PERSIST_BWD=1"; do
  F=$OUT/r02b_bench_$(echo $V | tr ' =' '__').json
  env $V timeout 120 $B > $F 2> ${F%.json}.err
  echo -n "   $V: "; line $F
done
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
export LSTM_NO_GRAPH=1
$CMD > $OUT/r02b_plain.log 2>&1 &&
ncu --cache-control none --clock-control none --metrics $M -k regex:k_fwd_step -s 900 -c 6 --csv --log-file $OUT/r02b_ncu_fwd_nocacheflush.csv $CMD > $OUT/r02b_ncu_fwd.log 2>&1
ncu --cache-control none --clock-control none --metrics $M -k regex:k_bwd_step -s 900 -c 6 --csv --log-file $OUT/r02b_ncu_bwd_nocacheflush.csv $CMD > $OUT/r02b_ncu_bwd.log 2>&1
LSTM_L2HINT=3 ncu --cache-control none --clock-control none --metrics $M -k regex:k_bwd_step -s 900 -c 6 --csv --log-file $OUT/r02b_ncu_bwd_hint3.csv $CMD > $OUT/r02b_ncu_bwd_h3.log 2>&1
tail -n 40 $OUT/r02b_ncu_fwd_nocacheflush.csv | cut -c1-400
