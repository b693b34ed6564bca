"""Where one timestep of the EXPERIMENTAL persistent forward recurrence (tc_persist.cu) spends its time: clock64 stamps of
CTA (0,0) around timestep 4.  Run on a B200:  LSTM_TC_DEBUG=1 LSTM_PERSIST_FWD=1 python scripts/persist_clocks.py [cfg4]
(add LSTM_PERSIST_SPREAD=1 to compare the one-line-per-counter grid barrier)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LSTM_TC_DEBUG", "1")
os.environ.setdefault("LSTM_PERSIST_FWD", "1")
import bench  # noqa: E402
import eigen_lstm_b200 as el  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
cfg = bench.WORKLOADS[wl]
N, B, S = cfg["N"], cfg["B"], cfg["S"]
g = el.LSTM(256, N, S, B, dtype=el.BF16)
g.init_params(0, 0.01, 1.0)
text = bench.synthetic_text(B * (S * 8) + 1000)
g.load_text(text.tobytes())
g.set_positions([S + b * S * 8 for b in range(B)])
g.train_text(3, stride=S - 1, lr=0.001, want_losses=False)
out = np.zeros(32, dtype=np.int64)
assert g.lib.lstm_debug_kernel_clocks(g.ctx, out.ctypes.data_as(C.c_void_p)) == 0
d = out[:9] - out[0]
names = ["producer at grid barrier", "barrier passed", "first stage landed", "last MMA issued", "accumulator complete",
         "accumulator in smem", "h(t) announced", "stores done", "producer at NEXT grid barrier"]
for n, v in zip(names, d):
    print(f"{v:8d} cycles  {n}")
print(f"timestep period: {d[8]} cycles = {d[8] / 1.965e3:.2f} us at 1965 MHz "
      "(note: the forward-step stamps of the per-step kernel share slots [0..8]; run with LSTM_PERSIST_FWD=1 only)")
