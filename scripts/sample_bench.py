"""Config 5 of BASELINE.json: latency-bound sampling — generate n characters at batch 1 from an H=1024 model on the
persistent recurrent kernel.  Prints one JSON line (us per character, chars/s).  Run on one B200."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eigen_lstm_b200 as el  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
g = el.LSTM(256, N, 2, 1)
g.init_params(seed=0, std=0.05)
g.sample(2000, seed=1)                       # warm-up
t0 = time.perf_counter()
out = g.sample(n, seed=0)
dt = time.perf_counter() - t0
t1 = time.perf_counter()
text = bytes(out[: min(n, 200000)])
bpc = g.test(text)
dt_eval = time.perf_counter() - t1
print(json.dumps({"workload": f"cfg5: sample {n} chars at batch 1, H={N} (weights N(0,0.05), seed 0)", "us_per_char": dt / n * 1e6,
                  "chars_per_s": n / dt, "seconds": dt, "distinct_bytes": len(set(out.tolist())),
                  "eval_us_per_char": dt_eval / len(text) * 1e6, "eval_bpc_of_own_sample": bpc,
                  "kernel": "k_recur_persist (cooperative, 128 CTAs, U rows resident in shared memory, 2 grid barriers per char "
                            "for sampling, 1 for evaluation)"}))
