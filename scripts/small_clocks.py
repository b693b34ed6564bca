"""In-kernel SM-clock stamps of the one-kernel training path (train_small.cu) at the reference's default shape.
Run on the GPU box:  LSTM_TC_DEBUG=1 python scripts/small_clocks.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["LSTM_TC_DEBUG"] = "1"
import ctypes as C
import numpy as np
import eigen_lstm_b200 as el

N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = el.LSTM(256, N, S, 1)
g.init_params(seed=0, std=0.01)
text = np.random.default_rng(0).integers(32, 127, 200000, dtype=np.uint8).tobytes()
g.load_text(text)
g.train_text(2000, stride=1, lr=0.1, want_losses=False)
g.sync()
out = (C.c_longlong * 32)()
assert g.lib.lstm_debug_kernel_clocks(g.ctx, out) == 0
a = np.array(out[:16]); b = np.array(out[16:])
t0 = a[0]
print(f"N={N} S={S}: one iteration = {a[8] - a[0]} SM cycles")
print("group A (thread 0):")
for name, i, j in [("window prefetch + forward", 0, 2), ("softmax (group B) + barrier", 2, 3), ("dHy", 3, 4), ("bptt", 4, 5),
                   ("U, b, W columns: Adagrad", 5, 6), ("wait for group B", 6, 8)]:
    print(f"  {name:30s} +{a[j] - a[i]:7d}   @ {a[j] - t0}")
print("group B (thread 256):")
for name, i in [("logits+softmax+loss", 4), ("waited for dHy", 5), ("Why + by: Adagrad", 6)]:
    print(f"  {name:30s} @ {b[i] - t0}")
