"""Small end-to-end run that touches every hand-written kernel family, for `compute-sanitizer --tool memcheck python scripts/sanitizer_run.py`:
persistent pair recurrences (N=2048, B=256), persistent single-tile recurrences (N=1024, B=128), the per-timestep kernels (B=384),
the weight-gradient / logits GEMMs, Adagrad + operand refresh, the fp32 SIMT path, the window pipeline and the batch-1 persistent
sampling / evaluation kernel."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eigen_lstm_b200 as el  # noqa: E402

text = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "enwik6_head.bin"), "rb").read()
for (N, S, B, dt) in ((2048, 4, 256, el.BF16), (1024, 4, 128, el.BF16), (512, 3, 384, el.BF16), (64, 5, 3, el.F32)):
    g = el.LSTM(256, N, S, B, dtype=dt)
    g.init_params(1, 0.02, 1.0)
    g.load_text(text)
    g.set_positions([S + 150 * b for b in range(B)])
    losses = g.train_text(3, stride=S - 1, lr=0.01)
    assert np.all(np.isfinite(losses)), losses
    print("trained", N, S, B, "bf16" if dt else "f32", g.variant() if dt else "", losses.round(3).tolist(), flush=True)
    g.close()
g = el.LSTM(256, 256, 3, 1)
g.init_params(2, 0.05)
print("sampled", bytes(g.sample(64, seed=3))[:16], "bpc", round(g.test(text[:600]), 4), flush=True)
print("sanitizer_run: done")
