#!/bin/bash
# round 2, call d: the persistent pair/split-K BPTT recurrence (tc_recur.cu): parity, bench A/B, clock stamps
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
timeout 600 python -m pytest tests/test_gpu_bench_variants.py tests/test_gpu_options.py tests/test_gpu_parity_bf16.py -x -q 2>&1 | tail -15 | tee $OUT/r02d_pytest.txt
for V in "LSTM_BWD_RECUR=-1" "LSTM_BWD_RECUR=0" "LSTM_BWD_RECUR=128"; do
  F=$OUT/r02d_bench_$(echo $V | tr ' =' '__').json
  env $V timeout 120 $B > $F 2> ${F%.json}.err
  echo -n "   $V: "; line $F; tail -2 ${F%.json}.err
done
LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -26 | tee $OUT/r02d_clocks_256.txt
LSTM_BWD_RECUR=128 LSTM_TC_DEBUG=1 timeout 100 python scripts/recur_clocks.py cfg4 2>&1 | tail -13 | tee $OUT/r02d_clocks_128.txt
