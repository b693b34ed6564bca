"""The reference's default shape (batch 1, N = 64; R/lstm.cc:52-58) trains inside ONE persistent kernel
(eigen_lstm_b200/csrc/train_small.cu).  That kernel accumulates every contraction in the same order as the general
launch-per-kernel fp32 path and shares its scalar functions, so the two must agree BIT FOR BIT — losses, parameters,
Adagrad memory, carried state, and the gradients / activations left behind by the last iteration.  (The general path
itself is checked against the CPU oracle in test_gpu_parity_f32.py, whose config-1 tests now run through this kernel.)"""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _make(N, S, seed, opts, text=None):
    import eigen_lstm_b200 as el
    M, B = 256, 1
    g = el.LSTM(M, N, S, B)
    g.set_params(orc.init_params(M, N, seed=seed, sd=0.05, forget_bias=opts.get("forget_bias", 0.0)))
    rng = np.random.default_rng(seed)
    g.set_state(rng.normal(0, 0.1, (N, B)), rng.normal(0, 0.1, (N, B)))
    if "clip" in opts:
        g.set_clip(opts["clip"])
    if "loss" in opts:
        g.set_loss_mode(opts["loss"])
    if "shift" in opts:
        g.set_softmax_shift(opts["shift"])
    if text is not None:
        g.load_text(text)
    return g


CASES = [
    (64, 3, 1, 0.1, {}),                                   # the reference's defaults (T = 2)
    (64, 3, 2, 0.1, {"forget_bias": 1.0}),                 # stride = T: non-overlapping windows
    (64, 2, 1, 0.1, {}),                                   # T = 1 in the T <= 2 instantiation (zero-padded second term)
    (64, 5, 1, 0.05, {"clip": 0.01}),                      # T = 4: the longest window the kernel takes; clipping active
    (64, 4, 2, 0.05, {"shift": "global", "loss": "last-ln"}),   # T = 3 in the T <= 4 instantiation; class_batch's loss and shift
    (64, 5, 4, 0.3, {}),                                   # large steps: Adagrad memory crosses the m >= 2^-9 shortcut early
]


@pytest.mark.parametrize("N,S,stride,lr,opts", CASES)
def test_one_kernel_training_equals_the_launch_per_kernel_path(alice, N, S, stride, lr, opts):
    iters = 400
    a = _make(N, S, 5, opts, alice)
    assert a.variant()["train_small"] == 1
    l0 = a.launch_count()
    la = a.train_text(iters, stride=stride, lr=lr)
    assert a.launch_count() - l0 == 1                      # all iterations in ONE launch
    b = _make(N, S, 5, opts, alice)
    b.set_profiling(True)                                  # per-phase events: the general path, plain stream launches
    lb = b.train_text(iters, stride=stride, lr=lr)
    assert b.launch_count() > iters
    assert np.all(np.isfinite(la))
    assert np.array_equal(la, lb)
    for w, (p, q) in enumerate(zip(a.params(), b.params())):
        assert np.array_equal(p, q), orc.NAMES[w]
    for w, (p, q) in enumerate(zip(a.adagrad_mem(), b.adagrad_mem())):
        assert np.array_equal(p, q), orc.NAMES[w]
    for w, (p, q) in enumerate(zip(a.grads(), b.grads())):   # of the last iteration
        assert np.array_equal(p, q), orc.NAMES[w]
    for p, q in zip(a.get_state(), b.get_state()):
        assert np.array_equal(p, q)
    for what in ("h", "c", "g", "probs", "dhy", "dg"):
        for t in range(1, S):
            assert np.array_equal(a.activation(what, t), b.activation(what, t)), (what, t)
    xa, ta = a.window(); xb, tb = b.window()
    assert np.array_equal(xa, xb) and np.array_equal(ta, tb)
    assert np.array_equal(a.positions(), b.positions())
    # and the two keep agreeing when the call is split (state, Adagrad memory and the event counter survive the launch)
    la2 = np.concatenate([a.train_text(7, stride=stride, lr=lr), a.train_text(1, stride=stride, lr=lr)])
    lb2 = b.train_text(8, stride=stride, lr=lr)
    assert np.array_equal(la2, lb2)


def test_host_windows_step_by_step(alice):
    """lstm_train_step (one iteration per launch, window from the host) == the piecewise calls of the general path."""
    N, S, stride, lr = 64, 3, 1, 0.1
    a = _make(N, S, 11, {"clip": 0.5})
    b = _make(N, S, 11, {"clip": 0.5})
    o = orc.Oracle(256, N, S, 1, "f32")
    for i in range(100):
        o.advance(alice, stride)
        x, t = o.window()
        la = a.train_step(x, t, stride=stride, lr=lr)
        lb = b.forward(x, t); b.backward(); b.adagrad(lr, 1e-10, 0.5); b.carry(stride)
        assert la == lb, i
    for p, q in zip(a.params(), b.params()):
        assert np.array_equal(p, q)


def test_shapes_outside_the_kernel_use_the_general_path():
    import eigen_lstm_b200 as el
    for (N, S, B) in [(64, 3, 2), (32, 3, 1), (64, 6, 1), (128, 3, 1)]:
        assert el.LSTM(256, N, S, B).variant()["train_small"] == 0
    assert el.LSTM(256, 64, 3, 1, dtype=el.BF16).variant()["train_small"] == 0
