"""The kernel instantiations the BENCHMARK runs, against the oracle.

`pick_bn` (csrc/tc_path.cu) chooses the tile width from the number of CTAs a shape yields, so the small shapes of
tests/test_gpu_parity_bf16.py all resolve to the narrowest instantiation.  Here the hidden size and batch are those of
BASELINE.json's configurations (window shortened to S = 3 so the CPU side finishes in seconds), the test ASSERTS which
variant the library selected (lstm_debug_variant), and every intermediate and all five gradients are compared —
the reference's own practice of running both implementations in lock-step and diffing every matrix
(OV/lstm_eigen_CUDA/lstm.cu:416-497,564-648).  CPU side: oracle/oracle_blas.py (the lstm_eigen_BLAS structure on
numpy's BLAS, itself checked against the C++ oracle in tests/test_oracle_blas.py)."""
import numpy as np
import pytest

from oracle import oracle as orc
from oracle import oracle_blas as ob
from tests.util import random_window

pytestmark = pytest.mark.gpu


def frob_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# (N, S, B, expected variant): config 4, the BN = 64 middle case, config 3, config 2
CASES = [
    (2048, 3, 256, dict(fwd_bn=128, fwd_pair=1, wgrad_bn=256, bwd_persistent=128, fwd_persistent=128)),
    (2048, 4, 200, dict(fwd_bn=128, fwd_pair=1, wgrad_bn=256, bwd_persistent=128, fwd_persistent=128)),   # padding rows in the second batch tile
    (1024, 3, 256, dict(wgrad_bn=256, fwd_persistent=128, bwd_persistent=128)),
    (1024, 4, 128, dict(wgrad_bn=256, fwd_persistent=32, bwd_persistent=32)),     # config 3: one batch tile, U resident
    (512, 4, 64, dict(wgrad_bn=128, fwd_persistent=32, bwd_persistent=32)),       # config 2
    (2048, 3, 100, dict(wgrad_bn=256, fwd_persistent=64, bwd_persistent=64)),
    (1024, 3, 384, dict(fwd_bn=128, fwd_pair=0, wgrad_bn=256, fwd_persistent=0, bwd_persistent=0)),   # three batch tiles: per-timestep kernels
    (1024, 3, 512, dict(fwd_bn=128, fwd_pair=1, wgrad_bn=256, fwd_persistent=0, bwd_persistent=0)),
    (512, 3, 512, dict(fwd_bn=64, fwd_pair=1, bwd_bn=64, wgrad_bn=128, fwd_persistent=0, bwd_persistent=0)),   # ... as cta_group::2 pairs
]


def _pair(M, N, S, B, seed):
    import eigen_lstm_b200 as el
    sd = 0.5 / np.sqrt(N)                      # pre-activations of O(1) whatever the width
    params = orc.init_params(M, N, seed=seed, sd=sd, forget_bias=0.5)
    rng = np.random.default_rng(seed + 100)
    params[2] = (params[2] + rng.normal(0, 0.1, params[2].shape)).astype(np.float32)
    params[4] = rng.normal(0, 0.1, params[4].shape).astype(np.float32)
    q = ob.BlasOracle(M, N, S, B)
    q.set_params(params)
    g = el.LSTM(M, N, S, B, dtype=el.BF16)
    g.set_params(params)
    h0 = rng.normal(0, 0.3, (N, B)).astype(np.float32)
    c0 = rng.normal(0, 0.3, (N, B)).astype(np.float32)
    q.h[0][...] = h0; q.c[0][...] = c0
    g.set_state(h0, c0)
    return q, g


@pytest.mark.parametrize("N,S,B,want", CASES)
def test_benchmark_instantiations_vs_oracle(N, S, B, want):
    M = 256
    q, g = _pair(M, N, S, B, seed=3)
    v = g.variant()
    for k, val in want.items():
        assert v[k] == val, (k, v)
    rng = np.random.default_rng(5)
    x, t = random_window(rng, M, S, B)
    q.set_window(x, t)
    lq = q.forward()
    lg = g.forward(x, t)
    assert abs(lg - lq) <= 2e-2 * abs(lq), (lg, lq)
    for tt in range(1, S):
        assert np.max(np.abs(g.activation("g", tt) - q.g[tt])) < 4e-2, ("g", tt)
        assert np.max(np.abs(g.activation("c", tt) - q.c[tt])) < 4e-2, ("c", tt)
        assert np.max(np.abs(g.activation("h", tt) - q.h[tt])) < 4e-2, ("h", tt)
        assert np.max(np.abs(g.activation("probs", tt) - q.probs[tt])) < 2e-2, ("probs", tt)
    q.backward()
    g.backward()
    for tt in range(S - 1, 0, -1):
        assert frob_rel(g.activation("dg", tt), q.dg[tt]) < 6e-2, ("dg", tt)
    for name, a, b in zip(orc.NAMES, g.grads(), q.grads):
        assert frob_rel(a, b) < 6e-2, name


def test_benchmark_shape_training_iterations_follow_the_oracle():
    """Five full iterations (forward, BPTT, Adagrad, carry) at config 4's width and batch: graph capture + replay included."""
    M, N, S, B = 256, 2048, 3, 256
    q, g = _pair(M, N, S, B, seed=9)
    rng = np.random.default_rng(1)
    for it in range(5):
        x, t = random_window(rng, M, S, B, nulls=False)
        q.set_window(x, t)
        lq = q.forward(); q.backward(); q.adagrad(0.002); q.carry(S - 1)
        lg = g.train_step(x, t, stride=S - 1, lr=0.002)
        assert abs(lg - lq) <= 3e-2 * abs(lq), (it, lg, lq)
    for name, a, b in zip(orc.NAMES, g.params(), q.params()):
        assert frob_rel(a, b) < 3e-2, name


@pytest.mark.parametrize("N,S,B,want_persistent", [(2048, 13, 256, 128), (1024, 13, 128, 32), (512, 9, 64, 32)])
def test_persistent_recurrences_agree_with_the_per_timestep_kernels(N, S, B, want_persistent):
    """The persistent recurrences (tc_recur.cu) against the launch-per-timestep kernels (LSTM_*_RECUR=0) on the same
    window over S-1 timesteps at the widths and batches of configs 4, 3 and 2: both contract the same bf16 operands, so only the
    fp32 summation order of the split-K partials differs."""
    import os
    import subprocess
    import sys
    import tempfile
    code = r"""
import sys, numpy as np
import eigen_lstm_b200 as el
M, N, S, B = 256, int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
g = el.LSTM(M, N, S, B, dtype=el.BF16)
g.init_params(3, 0.01, 1.0)
rng = np.random.default_rng(0)
g.set_state(rng.normal(0, 0.3, (N, B)).astype(np.float32), rng.normal(0, 0.3, (N, B)).astype(np.float32))
x = rng.integers(0, M, (S, B)).astype(np.int32); t = rng.integers(0, M, (S, B)).astype(np.int32)
loss = g.forward(x, t); g.backward()
v = g.variant()
np.savez(sys.argv[1], loss=loss, v=np.array([v["fwd_persistent"], v["bwd_persistent"]]), dg1=g.activation("dg", 1), dg7=g.activation("dg", 7),
         h5=g.activation("h", 5), **{n: a for n, a in zip(["W", "U", "b", "Why", "by"], g.grads())})
"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    envs = [{}, {"LSTM_BWD_RECUR": "0", "LSTM_FWD_RECUR": "0"}]
    if B > 128:
        envs.append({"LSTM_BWD_RECUR": "256"})
    out = []
    with tempfile.TemporaryDirectory() as d:
        for i, env in enumerate(envs):
            path = os.path.join(d, f"r{i}.npz")
            subprocess.run([sys.executable, "-c", code, path, str(N), str(S), str(B)], check=True, cwd=root, env=dict(os.environ, **env),
                           timeout=300)
            out.append(dict(np.load(path)))
    a, b = out[0], out[1]
    assert list(a["v"]) == [want_persistent, want_persistent] and list(b["v"]) == [0, 0]
    assert abs(a["loss"] - b["loss"]) < 1e-5 * abs(b["loss"])   # persistent vs per-timestep forward: same bf16 operands, other fp32 order
    assert frob_rel(a["h5"], b["h5"]) < 2e-3
    for k in ("dg1", "dg7", "W", "U", "b", "Why", "by"):
        assert frob_rel(a[k], b[k]) < 3e-3, k             # dg is rounded to bf16 every timestep: summation-order noise x 12 steps
    if B > 128:
        c = out[2]
        assert int(c["v"][1]) == 256 and a["loss"] == c["loss"]   # same forward kernel
        for k in ("dg1", "dg7", "W", "U", "b", "Why", "by"):
            assert frob_rel(c[k], b[k]) < 3e-3, k


@pytest.mark.parametrize("N,S,B", [(2048, 9, 256), (1024, 9, 128)])
def test_kernels_beside_the_recurrences_change_nothing(N, S, B):
    """K3 follows the running forward recurrence timestep by timestep and K6c runs beside the BPTT recurrence (programmatic
    dependent launches on the SMs the persistent kernels leave free).  Per-phase profiling keeps the kernels one after the other:
    both orders must give the same bits — losses, gradients, parameters — over several replayed iterations."""
    import eigen_lstm_b200 as el
    M = 256
    rng = np.random.default_rng(5)
    text = rng.integers(32, 127, 40000, dtype=np.uint8).tobytes()
    pos = [S + 97 * b for b in range(B)]
    out = []
    for profiling in (False, True):
        g = el.LSTM(M, N, S, B, dtype=el.BF16)
        assert g.variant()["fwd_persistent"] and g.variant()["bwd_persistent"]
        g.init_params(7, 0.01, 1.0)
        g.load_text(text); g.set_positions(pos)
        g.set_profiling(profiling)
        losses = g.train_text(6, stride=S - 1, lr=0.01)
        out.append((losses, g.params(), g.grads()))
        g.close()
    (la, pa, ga), (lb, pb, gb) = out
    assert np.all(np.isfinite(la)) and np.array_equal(la, lb)
    for name, a, b in zip(orc.NAMES, ga, gb):
        assert np.array_equal(a, b), name
    for name, a, b in zip(orc.NAMES, pa, pb):
        assert np.array_equal(a, b), name
