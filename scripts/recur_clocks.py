"""SM-clock stamps taken inside the persistent recurrence kernels (CTA 0, one timestep in the middle of the window).
Run on a B200:  LSTM_TC_DEBUG=1 python scripts/recur_clocks.py [cfg4]"""
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
print("variant:", g.variant())
g.init_params(0, 0.01, 1.0)
text = bench.synthetic_text(B * (S * 8) + 1000)
g.load_text(text.tobytes())
g.set_positions([S + b * S * 8 for b in range(B)])
g.train_text(3, stride=S - 1, lr=0.001, want_losses=False)
out = np.zeros(32, dtype=np.int64)
assert g.lib.lstm_debug_kernel_clocks(g.ctx, out.ctypes.data_as(C.c_void_p)) == 0
names = ["producer at grid barrier", "barrier passed", "first recurrent stage landed", "last MMA issued", "accumulator complete",
         "drained (slices stored)", "exchange complete / announced (fwd)", "dg(t) / h(t) released", "producer at NEXT grid barrier"]
for title, o in (("forward recurrence", 0), ("BPTT recurrence", 16)):
    d = out[o:o + 9] - out[o]
    print(title)
    for n, v in zip(names, d):
        print(f"  {v:8d} cycles  {n}")
    print(f"  timestep period: {d[8]} cycles = {d[8] / 1.965e3:.2f} us at 1965 MHz")
    if o == 16:
        x = out[o + 9:o + 13] - out[o]
        print(f"  BPTT detail: exchange release done {x[0]}, exchange poll done {x[1]}, partials summed {x[2]}, gate math + stores issued {x[3]}")
