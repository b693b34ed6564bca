"""North-star check: the bf16-GEMM / fp32-accumulate path reaches a final bits-per-char within 0.01 of the fp32 path
(same text, seed, initial weights, schedule).  Runs on one B200:  python scripts/bpc_bf16_vs_f32.py [N B S iters lr]

Text: the committed enwik6 (tests/golden/enwik6.txt, 10^6 bytes: BASELINE config 2's corpus and the stand-in for config 3's
missing enwik7); first 95 % for training with B streams, last 5 % held out (OV/lstm_eigen_class_batch/lstm.cc:55-59) and scored
with the reference's test() recipe (OV/lstm_eigen_class_CUDA/lstm.cc:661-720).  BPC_TEXT=head uses the 64 KiB head instead."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_lstm_b200 as el  # noqa: E402

N, B, S, iters, lr = 256, 32, 51, 1500, 0.02
if len(sys.argv) > 1:
    N, B, S, iters = [int(v) for v in sys.argv[1:5]]
    lr = float(sys.argv[5])
# BPC_MEM0=c: Adagrad memory initialised to c instead of 0 (an "initial accumulator value", same for every run).  With m = 0 the
# first update of EVERY weight is +-lr whatever the size of its gradient (d / sqrt(d^2)), i.e. a sign function of gradients that
# are mostly noise: a 1-ulp perturbation of the weights then moves the held-out bits/char by 0.1-0.3 after 3000 iterations
# (profiles/r02*_bpc_*lr*.json), which buries any bf16-vs-fp32 effect.  With m0 > 0 small gradients give small steps and the
# comparison measures precision instead of chaos.
mem0 = float(os.environ.get("BPC_MEM0", "0"))
head = os.environ.get("BPC_TEXT") == "head"
text = open(os.path.join(ROOT, "tests", "golden", "enwik6_head.bin" if head else "enwik6.txt"), "rb").read()
cut = len(text) * (90 if head else 95) // 100
train, held = text[:cut], text[cut:]
pos = [S + (len(train) - 2 * S) * b // B for b in range(B)]
res = {}
# "f32_ulp": the fp32 path again with every weight moved by ~1 ulp — the trajectory noise floor of this chaotic system
for name, dt in (("f32", el.F32), ("bf16", el.BF16), ("f32_ulp", el.F32)):
    g = el.LSTM(256, N, S, B, dtype=dt)
    g.init_params(seed=1, std=0.01, forget_bias=1.0)
    if name == "f32_ulp":
        rng = np.random.default_rng(0)
        for w in range(5):
            p = g.get(0, w)
            g.set(0, w, (p * (1.0 + 1.2e-7 * rng.choice([-1.0, 1.0], p.shape))).astype(np.float32))
    if mem0 > 0:
        for w in range(5):
            g.set(2, w, np.full(g.shape(w), mem0, dtype=np.float32))
    g.load_text(train)
    g.set_positions(pos)
    curve = []
    for k in range(iters // 100):
        l = g.train_text(100, stride=S - 1, lr=lr)
        curve.append(float(l.mean() / (S - 1)))
    res[name] = {"train_bpc_curve": curve, "heldout_bpc": g.test(held)}
    print(name, "train bpc (last 100 its)", curve[-1], "held-out bpc", res[name]["heldout_bpc"], flush=True)
res["abs_diff_heldout"] = abs(res["f32"]["heldout_bpc"] - res["bf16"]["heldout_bpc"])
res["abs_diff_train"] = abs(res["f32"]["train_bpc_curve"][-1] - res["bf16"]["train_bpc_curve"][-1])
res["noise_floor_heldout_f32_vs_f32_plus_1ulp"] = abs(res["f32"]["heldout_bpc"] - res["f32_ulp"]["heldout_bpc"])
res["config"] = dict(N=N, B=B, S=S, iters=iters, lr=lr, adagrad_mem0=mem0, text="enwik6 head 64 KiB, 90/10 split" if head else "enwik6 (10^6 bytes), 95/5 split",
                     variant=g.variant())
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", os.environ.get("BPC_OUT", "bpc_bf16_vs_f32.json")), "w"), indent=1)
print("held-out |bf16 - f32| =", res["abs_diff_heldout"], " train |diff| =", res["abs_diff_train"],
      " noise floor |f32 - f32(+1ulp)| =", res["noise_floor_heldout_f32_vs_f32_plus_1ulp"])
