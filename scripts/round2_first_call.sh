#!/bin/bash
# scripts/round2_first_call.sh — everything round 1 could not run any more, in ONE gpurun call (one B200, ~4 min):
#
#   gpurun --timeout 420 -- bash scripts/round2_first_call.sh
#
# Writes gpurun_out/r02a_*.  Each experiment is an environment switch on top of the default build; a failing one is
# reported and skipped, the default path is untouched.  What each answers is in DESIGN.md section 9.
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
line() {  # file -> "chars/s  ms/step  fwd us  bwd us  final loss  launches"
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
echo "== 0. full GPU test suite on the default build"
timeout 200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $OUT/r02a_pytest.txt
echo "== 1. default bench (the reference point of this box)"
timeout 120 $B > $OUT/r02a_bench_default.json 2> $OUT/r02a_bench_default.err; line $OUT/r02a_bench_default.json
echo "== 2. persistent forward recurrence: writer-side proxy fence variants (identical loss = still correct)"
for WF in 1 2 0; do
  LSTM_PERSIST_FWD=1 LSTM_PERSIST_WFENCE=$WF timeout 120 $B > $OUT/r02a_bench_persist_wf$WF.json 2> $OUT/r02a_bench_persist_wf$WF.err
  echo -n "   WFENCE=$WF: "; line $OUT/r02a_bench_persist_wf$WF.json
  LSTM_TC_DEBUG=1 LSTM_PERSIST_FWD=1 LSTM_PERSIST_WFENCE=$WF timeout 60 python scripts/persist_clocks.py > $OUT/r02a_persist_clocks_wf$WF.txt 2>&1
  tail -10 $OUT/r02a_persist_clocks_wf$WF.txt
done
echo "== 3. K5 as CTA pairs in clusters of 2 with the counter-ordered split-K exchange"
LSTM_BWD_PAIR=2 timeout 120 python -m pytest tests/test_gpu_parity_bf16.py -x -q 2>&1 | tail -3 | tee $OUT/r02a_pytest_bwdpair2.txt
LSTM_BWD_PAIR=2 timeout 120 $B > $OUT/r02a_bench_bwdpair2.json 2> $OUT/r02a_bench_bwdpair2.err; line $OUT/r02a_bench_bwdpair2.json
echo "== 3b. persistent BPTT recurrence (LSTM_PERSIST_BWD=1): parity first, then the bench (identical loss = still correct)"
LSTM_PERSIST_BWD=1 timeout 120 python -m pytest tests/test_gpu_parity_bf16.py -x -q 2>&1 | tail -3 | tee $OUT/r02a_pytest_persist_bwd.txt
LSTM_PERSIST_BWD=1 timeout 120 $B > $OUT/r02a_bench_persist_bwd.json 2> $OUT/r02a_bench_persist_bwd.err; line $OUT/r02a_bench_persist_bwd.json
LSTM_PERSIST_BWD=1 LSTM_PERSIST_FWD=1 timeout 120 $B > $OUT/r02a_bench_persist_both.json 2> $OUT/r02a_bench_persist_both.err; line $OUT/r02a_bench_persist_both.json
echo "== 4. grid barrier microbenchmark"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gridbar scripts/gridbar_bench.cu && timeout 30 /tmp/gridbar 128 2000 | tee $OUT/r02a_gridbar.txt
echo "== 5. regression checks the end of round 1 could not run (DESIGN section 9, item 5)"
timeout 120 python - <<'PY' 2>&1 | tail -8 | tee gpurun_out/r02a_regressions.txt
import os, tempfile
import numpy as np
import eigen_lstm_b200 as el
G = os.path.join("tests", "golden")
alice = open(os.path.join(G, "alice29_head.bin"), "rb").read()
enwik = open(os.path.join(G, "enwik6_head.bin"), "rb").read()
M, N, S, B = 256, 32, 4, 2
# (a) a second lstm_load_text after graphs were captured must not replay the old corpus pointer
a = el.LSTM(M, N, S, B); a.init_params(5); a.load_text(alice); a.set_positions([10, 3000]); a.train_text(6, 1, 0.1)
path = os.path.join(tempfile.mkdtemp(), "ck.bin"); a.save_bin(path)
a.load_text(enwik); a.set_positions([20, 5000]); la = a.train_text(6, 1, 0.1)
b = el.LSTM(M, N, S, B); b.load_text(enwik); b.load_bin(path); b.set_positions([20, 5000]); lb = b.train_text(6, 1, 0.1)
print("load_text twice == fresh context:", bool(np.array_equal(la, lb)))
# (b) binary checkpoint of a bf16 context carries the live h(0)
M, N, S, B = 256, 64, 6, 4
c = el.LSTM(M, N, S, B, dtype=el.BF16); c.init_params(3, 0.05, 1.0); c.load_text(enwik); c.set_positions([10, 900, 1800, 2700])
c.train_text(8, S - 1, 0.05); c.save_bin(path); lc = c.train_text(5, S - 1, 0.05)
d = el.LSTM(M, N, S, B, dtype=el.BF16); d.load_text(enwik); d.load_bin(path); ld = d.train_text(5, S - 1, 0.05)
print("bf16 save_bin/load_bin resume bit-exact:", bool(np.array_equal(lc, ld)))
PY
echo "== 6. K2 operand-sharing A/B (after the elect.sync fix): no pair / multicast clusters"
for V in "LSTM_PAIR=0" "LSTM_FWD_CN=2" "LSTM_FWD_CN=4" "LSTM_FWD_CN=4 LSTM_FWD_CM=2" "LSTM_FWD_CN=2 LSTM_FWD_CM=2" "LSTM_NO_L2PIN=1" "LSTM_NO_PDL=1"; do
  F=$OUT/r02a_bench_$(echo $V | tr ' =' '__').json
  env $V timeout 120 $B > $F 2> ${F%.json}.err
  echo -n "   $V: "; line $F
done
echo "== 7. in-kernel clock stamps of the default per-step kernels"
LSTM_TC_DEBUG=1 timeout 60 python scripts/kernel_clocks.py cfg4 2>&1 | tail -3 | tee $OUT/r02a_kernel_clocks.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee $OUT/r02a_smi.txt
python - <<'PY' | tee $OUT/r02a_devprops.txt
import torch
p = torch.cuda.get_device_properties(0)
print(p)
print("L2", p.L2_cache_size, "smem/SM", p.shared_memory_per_multiprocessor, "smem/block optin", p.shared_memory_per_block_optin)
PY
echo "== done; copy what is worth keeping from gpurun_out/r02a_* into profiles/"
