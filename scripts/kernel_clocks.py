"""Print where the time goes INSIDE one forward-step and one BPTT-step kernel (clock64 stamps of CTA 0).
Run on a B200:  LSTM_TC_DEBUG=1 python scripts/kernel_clocks.py [cfg4|cfg3|cfg2]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LSTM_TC_DEBUG", "1")
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
for name, o in (("fwd step", 0), ("bwd step", 16)):
    d = out[o:o + 9] - out[o]
    print(f"{wl} {name}: prologue {d[4]} | first stage landed {d[2]} | last TMA issued {d[1]} | last MMA issued {d[3]} | "
          f"accum complete {d[5]} | remap/reduce done {d[6]} | math+stores done {d[7]} | exit {d[8]}  (SM cycles from entry)")
