"""Shared helpers for the parity tests: build matching oracle / GPU models and compare them."""
import numpy as np

from oracle import oracle as orc


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def random_window(rng, M, S, B, nulls=True):
    x = rng.integers(0, M, (S, B)).astype(np.int32)
    t = rng.integers(0, M, (S, B)).astype(np.int32)
    if nulls and S > 2:
        x[1, 0] = -1
        t[1, B - 1] = -1
    return x, t


def make_pair(M, N, S, B, seed=0, sd=0.1, forget_bias=0.0, dtype=0, state_sd=0.1):
    """(oracle, gpu) with identical injected parameters and carried-in state."""
    import eigen_lstm_b200 as el
    params = orc.init_params(M, N, seed=seed, sd=sd, forget_bias=forget_bias)
    rng = np.random.default_rng(seed + 100)
    params[2] = (params[2] + rng.normal(0, sd, params[2].shape)).astype(np.float32)
    params[4] = rng.normal(0, sd, params[4].shape).astype(np.float32)
    o = orc.Oracle(M, N, S, B, "f32")
    g = el.LSTM(M, N, S, B, dtype=dtype)
    o.set_params(params)
    g.set_params(params)
    h0 = rng.normal(0, state_sd, (N, B)).astype(np.float32)
    c0 = rng.normal(0, state_sd, (N, B)).astype(np.float32)
    o.set_state("h", 0, h0)
    o.set_state("c", 0, c0)
    g.set_state(h0, c0)
    return o, g, params
