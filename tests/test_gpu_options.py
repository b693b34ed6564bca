"""Options of the training path (lstm_set_option): fused gradient clipping (north-star item 4), the class_batch snapshot's
loss report (last timestep, nats) and its global-max softmax shift."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import make_pair, random_window, rel_err

pytestmark = pytest.mark.gpu
M = 256


@pytest.mark.parametrize("dtype", [0, 1])
def test_clip_on_the_training_path_vs_oracle(dtype):
    """lstm_train_step with clip > 0 == oracle forward/backward, gradients clamped to [-clip, clip], Adagrad (R/lstm.cc:259-272)."""
    N, S, B = 64, 5, 6
    clip = 0.02
    o, g, _ = make_pair(M, N, S, B, seed=4, sd=0.1, dtype=dtype)
    g.set_clip(clip)
    rng = np.random.default_rng(8)
    hit = 0
    for it in range(4):
        x, t = random_window(rng, M, S, B, nulls=False)
        o.set_window(x, t)
        lo = o.forward(); o.backward()
        for w in range(5):
            d = o.get(orc.GRAD, w)
            hit += int(np.sum(np.abs(d) > clip))
            o.set(orc.GRAD, w, np.clip(d, -clip, clip))
        o.adagrad(0.05); o.carry(S - 1)
        lg = g.train_step(x, t, stride=S - 1, lr=0.05)
        assert abs(lg - lo) <= (1e-4 if dtype == 0 else 3e-2) * abs(lo), (it, lg, lo)
    assert hit > 100                                   # the clamp was active
    # fp32: the updated weights themselves.  bf16: Adagrad's first steps move every weight by +-lr whatever the gradient's size, so
    # a gradient that differs in sign at the 1e-3 level flips a weight by 2*lr; the accumulated squared CLIPPED gradients
    # (Adagrad memory) are the smooth quantity to compare, and they also show the clamp: no entry can exceed iterations * clip^2.
    for name, a, b in zip(orc.NAMES, g.params(), o.params()):
        if dtype == 0:
            assert rel_err(a, b) < 2e-4, name
    mem_g = g.adagrad_mem()
    mem_o = [o.get(orc.MEM, w) for w in range(5)]
    for name, a, b in zip(orc.NAMES, mem_g, mem_o):
        assert float(np.max(a)) <= 4 * clip * clip * (1 + 1e-5), name
        fro = float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b))
        assert fro < (1e-4 if dtype == 0 else 6e-2), (name, fro)
    # and with clip = 0 the very same call is the unclipped reference update again
    g.set_clip(0.0)
    x, t = random_window(rng, M, S, B, nulls=False)
    o.set_window(x, t); o.forward(); o.backward(); o.adagrad(0.05)
    g.train_step(x, t, stride=S - 1, lr=0.05)
    assert float(np.max(g.adagrad_mem()[1])) > 4 * clip * clip * 1.5   # unclipped gradients are larger than the clamp
    if dtype == 0:
        for name, a, b in zip(orc.NAMES, g.params(), o.params()):
            assert rel_err(a, b) < 2e-4, name


@pytest.mark.parametrize("dtype", [0, 1])
def test_last_timestep_natural_log_loss_report(dtype):
    """OV/lstm_eigen_class_batch/lstm.cc:308-319: the reported loss is the LAST timestep's sum_b -ln p[target]; here / B."""
    N, S, B = 64, 6, 5
    o, g, _ = make_pair(M, N, S, B, seed=2, sd=0.1, dtype=dtype)
    g.set_loss_mode("last-ln")
    rng = np.random.default_rng(3)
    x, t = random_window(rng, M, S, B, nulls=False)
    o.set_window(x, t); o.forward()
    p = o.state("probs", S - 1).astype(np.float64)
    want = float(np.sum(-np.log(p[t[S - 1], np.arange(B)]))) / B
    got = g.forward(x, t)
    assert abs(got - want) <= (2e-5 if dtype == 0 else 2e-2) * want, (got, want)
    # gradients still flow from every timestep (OV/lstm_eigen_class_batch/lstm.h:281-283)
    o.backward(); g.backward()
    for name, a, b in zip(orc.NAMES, g.grads(), o.grads()):
        assert rel_err(a, b) < (2e-4 if dtype == 0 else 6e-2), name
    g.set_loss_mode("log2-all")
    assert abs(g.forward(x, t) - o.forward()) <= (2e-5 if dtype == 0 else 2e-2) * want * 10


def test_global_max_softmax_shift_vs_oracle():
    """OV/lstm_eigen_class_batch/lstm.h:175 (y.array() -= y.maxCoeff() over the whole M x B matrix) on the fp32 path, with
    logits large enough that the unshifted exp would overflow float32."""
    N, S, B = 48, 5, 7
    o, g, params = make_pair(M, N, S, B, seed=6, sd=0.1, dtype=0)
    by = params[4].copy(); by[:, 0] += 95.0            # exp(95) overflows float32: only the shifted softmax is finite
    o.set(orc.PARAM, orc.BY, by); g.set(0, 4, by)
    o.set_options(softmax_shift=1)
    g.set_softmax_shift("global")
    rng = np.random.default_rng(1)
    x, t = random_window(rng, M, S, B, nulls=False)
    o.set_window(x, t)
    lo, lg = o.forward(), g.forward(x, t)
    assert np.isfinite(lo) and np.isfinite(lg) and abs(lg - lo) <= 1e-5 * abs(lo), (lg, lo)
    for tt in range(1, S):
        assert rel_err(g.activation("probs", tt), o.state("probs", tt)) < 1e-5
    o.backward(); g.backward()
    for name, a, b in zip(orc.NAMES, g.grads(), o.grads()):
        assert rel_err(a, b) < 2e-4, name
    g.set_softmax_shift("none")
    assert not np.isfinite(g.forward(x, t))            # the reference's own formula (R/lstm.cc:199-201) overflows here


def test_options_reject_bad_values():
    import eigen_lstm_b200 as el
    g = el.LSTM(M, 32, 3, 1)
    with pytest.raises(el.LstmError):
        g.set_clip(-1.0)
    with pytest.raises(el.LstmError):
        g._ck(g.lib.lstm_set_option(g.ctx, 99, 1.0))
