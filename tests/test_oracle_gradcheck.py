"""Backward of the oracle vs central differences, with the reference's own acceptance rule
(OV/lstm_eigen_class/lstm.cc:250-304: rel.err = |a-n|/|a+n|, fail if max > 1e-1 or mean > 1e-3;
delta = 1e-5, double precision, OV/lstm_eigen_class/lstm.h:131-170)."""
import numpy as np

from oracle import oracle as orc

LN2 = np.log(2.0)


def _setup(M=24, N=7, S=6, B=3, seed=0):
    rng = np.random.default_rng(seed)
    o = orc.Oracle(M, N, S, B, "f64")
    params = [rng.normal(0, 0.3, s) for s in [(4 * N, M), (4 * N, N), (4 * N, 1), (M, N), (M, 1)]]
    o.set_params(params)
    x = rng.integers(0, M, (S, B)).astype(np.int32)
    t = rng.integers(0, M, (S, B)).astype(np.int32)
    x[1, 0] = -1  # a not-yet-filled input column (warm-up, R/lstm.cc:84,170)
    t[1, 1] = -1  # a not-yet-filled target column
    o.set_window(x, t)
    o.set_state("h", 0, rng.normal(0, 0.1, (N, B)))
    o.set_state("c", 0, rng.normal(0, 0.1, (N, B)))
    return o, params


def test_reference_acceptance_rule_and_tight_bound():
    # full window (no -1 targets) so loss and gradient describe the same function
    o, params = _setup()
    x, t = o.window()
    t[1, 1] = 5
    o.set_window(x, t)
    o.forward(); o.backward()
    ana = o.grads()
    delta = 1e-5
    B = o.B
    mx, tot, cnt = 0.0, 0.0, 0
    for which in range(5):
        p = params[which].copy()
        for idx in np.ndindex(p.shape):
            q = p.copy(); q[idx] += delta; o.set(orc.PARAM, which, q); lp = o.forward()
            q = p.copy(); q[idx] -= delta; o.set(orc.PARAM, which, q); lm = o.forward()
            n = (lp - lm) * LN2 * B / (2 * delta)
            a = ana[which][idx]
            if abs(a) + abs(n) < 1e-5:
                continue
            r = abs(a - n) / abs(a + n)
            mx = max(mx, r); tot += r; cnt += 1
        o.set(orc.PARAM, which, p)
    assert mx < 1e-1 and tot / cnt < 1e-3      # the reference's rule
    assert mx < 1e-5 and tot / cnt < 1e-7, (mx, tot / cnt)   # what the restatement actually holds
