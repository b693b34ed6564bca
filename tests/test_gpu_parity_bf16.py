"""bf16 tensor-core path (tcgen05 / TMEM / TMA) vs the fp32 CPU oracle, through the C ABI.
Tolerances are those of bf16 operands (8 significand bits) with fp32 accumulation; a layout or
descriptor bug shows up as O(1) error, not O(1e-2)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import make_pair, random_window

pytestmark = pytest.mark.gpu


def frob_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


SHAPES = [  # (N, S, B)
    (64, 4, 3),       # one K block, one batch tile mostly padding
    (128, 6, 130),    # two batch tiles (Bp = 256), second almost empty
    (256, 3, 64),
    (512, 5, 32),     # cfg2's hidden size
]


@pytest.mark.parametrize("N,S,B", SHAPES)
def test_forward_backward_vs_oracle(N, S, B):
    import eigen_lstm_b200 as el
    M = 256
    o, g, _ = make_pair(M, N, S, B, seed=2, sd=0.08, dtype=el.BF16)
    rng = np.random.default_rng(11)
    x, t = random_window(rng, M, S, B)
    o.set_window(x, t)
    lo = o.forward()
    lg = g.forward(x, t)
    assert abs(lg - lo) <= 2e-2 * abs(lo), (lg, lo)
    for tt in range(1, S):
        assert np.max(np.abs(g.activation("g", tt) - o.state("g", tt))) < 3e-2, ("g", tt)
        assert np.max(np.abs(g.activation("c", tt) - o.state("c", tt))) < 3e-2, ("c", tt)
        assert np.max(np.abs(g.activation("h", tt) - o.state("h", tt))) < 3e-2, ("h", tt)
        assert np.max(np.abs(g.activation("probs", tt) - o.state("probs", tt))) < 2e-2, ("probs", tt)
    o.backward()
    g.backward()
    for tt in range(S - 1, 0, -1):
        assert frob_rel(g.activation("dg", tt), o.state("dg", tt)) < 6e-2, ("dg", tt)
    for name, a, b in zip(orc.NAMES, g.grads(), o.grads()):
        assert frob_rel(a, b) < 6e-2, name


def test_weight_gradient_gemm_exact_on_bf16_grid():
    """K6 (dW|dU|db and dWhy|dby as two GEMMs over all (t,b)) against a float64 product of the very
    bf16 operands the GPU stashed: only fp32 accumulation order differs."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 64, 5, 7
    o, g, _ = make_pair(M, N, S, B, seed=4, sd=0.1, dtype=el.BF16)
    rng = np.random.default_rng(3)
    x, t = random_window(rng, M, S, B, nulls=True)
    g.forward(x, t); g.backward()
    dU = np.zeros((4 * N, N)); dW = np.zeros((4 * N, M)); db = np.zeros((4 * N, 1))
    dWhy = np.zeros((M, N)); dby = np.zeros((M, 1))
    for tt in range(1, S):
        dg = g.activation("dg", tt).astype(np.float64)          # bf16 values
        hp = g.activation("h", tt - 1).astype(np.float64)
        ht = g.activation("h", tt).astype(np.float64)
        dy = g.activation("probs", tt).astype(np.float64)
        for b in range(B):
            if t[tt, b] >= 0:
                dy[t[tt, b], b] -= 1.0
        dU += dg @ hp.T
        db += dg.sum(1, keepdims=True)
        for b in range(B):
            if x[tt, b] >= 0:
                dW[:, x[tt, b]] += dg[:, b]
        dWhy += dy @ ht.T
        dby += dy.sum(1, keepdims=True)
    got = g.grads()
    for name, a, b in zip(orc.NAMES, got, [dW, dU, db, dWhy, dby]):
        assert frob_rel(a, b) < 2e-3, name   # probs are stored as bf16(p - onehot): the +1 round trip costs ~1e-3


def test_bf16_training_tracks_fp32(enwik6):
    """North-star: the bf16-GEMM path's loss follows the fp32 path (same text, seed, initial weights)."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 128, 17, 16
    params = orc.init_params(M, N, seed=5, sd=0.01, forget_bias=1.0)
    pos = [S + 1000 * b for b in range(B)]
    res = []
    for dt in (el.F32, el.BF16):
        g = el.LSTM(M, N, S, B, dtype=dt); g.set_params(params); g.load_text(enwik6); g.set_positions(pos)
        res.append(g.train_text(150, stride=S - 1, lr=0.02) / (S - 1))
    f32, b16 = res
    assert abs(f32[-20:].mean() - b16[-20:].mean()) < 0.05, (f32[-20:].mean(), b16[-20:].mean())
    assert b16[-20:].mean() < b16[:5].mean()     # it learns


def test_bf16_state_roundtrip_and_carry():
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 64, 4, 5
    g = el.LSTM(M, N, S, B, dtype=el.BF16); g.init_params(1, 0.05)
    rng = np.random.default_rng(0)
    h0 = rng.normal(0, 0.1, (N, B)).astype(np.float32); c0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    g.set_state(h0, c0)
    h, c = g.get_state()
    assert np.max(np.abs(h - h0)) < 1e-3 and np.array_equal(c, c0)
    x, t = random_window(rng, M, S, B, nulls=False)
    g.forward(x, t)
    h2 = g.activation("h", 2); c2 = g.activation("c", 2)
    g.carry(2)
    h, c = g.get_state()
    assert np.array_equal(h, h2) and np.array_equal(c, c2)


def test_bf16_rejects_unsupported_shapes():
    import eigen_lstm_b200 as el
    with pytest.raises(el.LstmError):
        el.LSTM(256, 20, 3, 1, dtype=el.BF16)     # N not a multiple of 64
    with pytest.raises(el.LstmError):
        el.LSTM(100, 64, 3, 1, dtype=el.BF16)     # M != 256
