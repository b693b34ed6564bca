"""GPU (fp32 SIMT path) vs the CPU oracle, through the C ABI.  Mirrors the reference's own
CPU<->GPU lock-step checks (OV/lstm_eigen_CUDA/lstm.cu:416-497,564-648; compare_lstm_states,
OV/lstm_eigen_class_CUDA/cu_lstm.h:398-415) with explicit thresholds."""
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import make_pair, random_window, rel_err

pytestmark = pytest.mark.gpu

SHAPES = [  # (M, N, S, B)
    (256, 64, 3, 1),     # R/lstm.cc defaults
    (256, 32, 5, 4),     # OV/lstm_eigen_class_batch defaults
    (256, 20, 7, 3),     # N not a multiple of the tile (OV/lstm_eigen_class: N = 20)
    (100, 72, 4, 70),    # B beyond one 64-row tile, M < 256
    (256, 128, 12, 16),
]


@pytest.mark.parametrize("M,N,S,B", SHAPES)
def test_forward_backward_every_intermediate(M, N, S, B):
    o, g, _ = make_pair(M, N, S, B, seed=1)
    rng = np.random.default_rng(5)
    x, t = random_window(rng, M, S, B)
    o.set_window(x, t)
    lo = o.forward()
    lg = g.forward(x, t)
    assert abs(lg - lo) <= 2e-6 * abs(lo), (lg, lo)
    for tt in range(1, S):
        for what in ("g", "c", "h", "probs"):
            assert rel_err(g.activation(what, tt), o.state(what, tt)) < 2e-5, (what, tt)
    o.backward()
    g.backward()
    for tt in range(1, S):
        assert rel_err(g.activation("dg", tt), o.state("dg", tt)) < 2e-5, ("dg", tt)
    for name, a, b in zip(orc.NAMES, g.grads(), o.grads()):
        assert rel_err(a, b) < 2e-5, name
    o.adagrad(0.1)
    g.adagrad(0.1)
    for name, a, b in zip(orc.NAMES, g.params(), o.params()):
        assert rel_err(a, b) < 2e-3, name   # the first Adagrad step is d/sqrt(d*d+eps): steep in d near 0 (kernel itself: bit-exact test below)
    for name, a, b in zip(orc.NAMES, g.adagrad_mem(), [o.get(orc.MEM, i) for i in range(5)]):
        assert rel_err(a, b) < 4e-5, name


def test_adagrad_kernel_bit_exact():
    """Same gradients in, same update out: m += d*d ; p -= lr*d/sqrtf(m+eps) (R/lstm.cc:259-272)."""
    M, N, S, B = 256, 24, 3, 2
    o, g, params = make_pair(M, N, S, B, seed=3)
    rng = np.random.default_rng(0)
    for it in range(3):
        for w in range(5):
            d = rng.normal(0, 10.0 ** rng.integers(-6, 1), params[w].shape).astype(np.float32)
            o.set(orc.GRAD, w, d)
            g.set(1, w, d)
        o.adagrad(0.1)
        g.adagrad(0.1)
        for a, b in zip(g.params(), o.params()):
            assert np.array_equal(a, b)
        for w in range(5):
            assert np.array_equal(g.get(2, w), o.get(orc.MEM, w))


def _cfg1_setup(alice):
    M, N, S, B = 256, 64, 3, 1
    params = orc.init_params(M, N, seed=1234, sd=0.01)
    h0 = orc.randn(N, S, 0, 0.1, 77)[:, 1:2]   # column that becomes h(0) after the first shift (R/lstm.cc:146,163)
    c0 = orc.randn(N, S, 0, 0.1, 78)[:, 1:2]
    return M, N, S, B, params, h0, c0


def test_cfg1_loss_trace_1000_iterations_resynchronised(alice):
    """North-star bar: per-step loss within 1e-4 relative of the Eigen path over the first 1000
    iterations (R/lstm.cc defaults N=64, M=256, S=3, B=1, lr=0.1, alice29.txt, same seed and weights).

    The reference's training trajectory at these defaults is CHAOTIC: the oracle itself, rebuilt with
    FMA contraction, leaves the 1e-4 band after ~60 iterations, and its float64 build after ~11
    (test_cfg1_free_running_divergence_is_intrinsic).  No implementation with a different rounding
    order can free-run inside 1e-4 for 1000 iterations, so the bar is applied the only way it is
    well-posed: every iteration starts from the oracle's exact state (weights, Adagrad memory, h, c)
    and must reproduce that iteration's loss within 1e-4 and its updated weights within 1e-4."""
    import eigen_lstm_b200 as el
    M, N, S, B, params, h0, c0 = _cfg1_setup(alice)
    o = orc.Oracle(M, N, S, B, "f32")
    o.set_params(params); o.set_state("h", 0, h0); o.set_state("c", 0, c0)
    g = el.LSTM(M, N, S, B)
    worst_loss, worst_par = 0.0, 0.0
    for i in range(1000):
        if i > 0:
            o.carry(1)
        o.advance(alice, 1)
        x, t = o.window()
        # resynchronise the GPU context from the oracle
        for w in range(5):
            g.set(0, w, o.get(orc.PARAM, w)); g.set(2, w, o.get(orc.MEM, w))
        g.set_state(o.state("h", 0), o.state("c", 0))
        ref = o.forward(); o.backward(); o.adagrad(0.1)
        got = g.train_step(x, t, stride=1, lr=0.1)
        if ref > 0:
            worst_loss = max(worst_loss, abs(got - ref) / ref)
        else:
            assert got == 0
        if i % 50 == 49:
            for a, b in zip(g.params(), o.params()):
                worst_par = max(worst_par, rel_err(a, b))
            hg, cg = g.get_state()
            worst_par = max(worst_par, rel_err(hg, o.state("h", 1)), rel_err(cg, o.state("c", 1)))
    assert worst_loss <= 1e-4, worst_loss
    assert worst_par <= 1e-4, worst_par


def test_cfg1_free_running_divergence_is_intrinsic(alice):
    """Free-running for 1000 iterations: the GPU path leaves the 1e-4 band no earlier than the
    reference algorithm's own rounding sensitivity (oracle vs the same oracle built with FMA), and the
    two trajectories stay statistically indistinguishable (mean loss of the last 200 iterations)."""
    import eigen_lstm_b200 as el
    M, N, S, B, params, h0, c0 = _cfg1_setup(alice)

    def cpu(mt):
        o = orc.Oracle(M, N, S, B, "f32", mt=mt)
        o.set_params(params); o.set_state("h", 1, h0); o.set_state("c", 1, c0)   # its first carry moves slot 1 -> 0
        return o.train(alice, 1000, stride=1, lr=0.1)[0]

    ref, ref_fma = cpu(False), cpu(True)
    o64 = orc.Oracle(M, N, S, B, "f64")
    o64.set_params(params); o64.set_state("h", 1, h0); o64.set_state("c", 1, c0)
    ref_f64 = o64.train(alice, 1000, stride=1, lr=0.1)[0]
    g = el.LSTM(M, N, S, B)
    g.set_params(params); g.set_state(h0, c0); g.load_text(alice)
    got = g.train_text(1000, stride=1, lr=0.1)
    nz = ref > 0

    def first_exceed(x):
        e = np.abs(x[nz] - ref[nz]) / ref[nz]
        return int(np.argmax(e > 1e-4)) if (e > 1e-4).any() else len(e)

    i_gpu, i_cpu = first_exceed(got), min(first_exceed(ref_fma), first_exceed(ref_f64))
    print(f"first iteration outside 1e-4: gpu {i_gpu}, oracle-with-FMA {first_exceed(ref_fma)}, oracle-f64 {first_exceed(ref_f64)}")
    assert i_gpu >= min(5, i_cpu // 2), (i_gpu, i_cpu)
    e = np.abs(got[nz] - ref[nz]) / ref[nz]
    assert e[:i_gpu].max() <= 1e-4
    assert abs(got[-200:].mean() - ref[-200:].mean()) < 0.15 * ref[-200:].mean()


def test_cfg1_free_running_1000_iterations_at_a_non_chaotic_learning_rate(alice):
    """The same 1000 FREE-RUNNING iterations at the reference's shape (N=64, S=3, B=1, alice29) with lr = 1e-3 instead of 0.1:
    Adagrad's first steps then move a weight by 1e-3 against an N(0, 0.01) init, the trajectory is no longer chaotic, and the
    GPU path must stay within 1e-4 relative of the oracle's per-iteration loss with NO resynchronisation — which rules out a slow
    systematic drift that the resynchronised test above could not see."""
    import eigen_lstm_b200 as el
    M, N, S, B, params, h0, c0 = _cfg1_setup(alice)
    o = orc.Oracle(M, N, S, B, "f32")
    o.set_params(params); o.set_state("h", 1, h0); o.set_state("c", 1, c0)   # its first carry moves slot 1 -> 0
    ref, _ = o.train(alice, 1000, stride=1, lr=1e-3)
    g = el.LSTM(M, N, S, B)
    g.set_params(params); g.set_state(h0, c0); g.load_text(alice)
    got = g.train_text(1000, stride=1, lr=1e-3)
    rel = np.abs(got - ref) / np.abs(ref)
    assert rel.max() < 1e-4, (int(rel.argmax()), float(rel.max()))
    for name, a, b in zip(orc.NAMES, g.params(), o.params()):
        assert rel_err(a, b) < 1e-4, name


def test_cfg1_device_pipeline_equals_host_windows(alice):
    """lstm_train_text (windows built by the K10 kernel) == lstm_train_step with host-built windows."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 64, 3, 1
    params = orc.init_params(M, N, seed=9, sd=0.01)
    a = el.LSTM(M, N, S, B); a.set_params(params); a.load_text(alice)
    b = el.LSTM(M, N, S, B); b.set_params(params)
    la = a.train_text(200, stride=1, lr=0.1)
    og = orc.Oracle(M, N, S, B, "f32")
    lb = np.zeros(200)
    for i in range(200):
        og.advance(alice, 1)
        x, t = og.window()
        lb[i] = b.train_step(x, t, stride=1, lr=0.1)
    assert np.array_equal(la, lb)
    for p, q in zip(a.params(), b.params()):
        assert np.array_equal(p, q)


@pytest.mark.parametrize("stride", [1, 9])
def test_batched_trajectory(enwik6, stride):
    """cfg2-style (batched, forget bias 1, random stream positions) at a size the oracle finishes in seconds."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 48, 10, 8
    params = orc.init_params(M, N, seed=21, sd=0.01, forget_bias=1.0)
    rng = np.random.default_rng(2)
    pos = rng.integers(S, len(enwik6), B)
    pos[0] = len(enwik6) - 5  # wraps to S almost immediately (OV/lstm_eigen_opt/lstm.cc:196-197)
    o = orc.Oracle(M, N, S, B, "f32"); o.set_params(params); o.set_positions(pos)
    g = el.LSTM(M, N, S, B); g.set_params(params); g.load_text(enwik6); g.set_positions(pos)
    iters = 60
    lr = 0.002   # small enough that 60 free-running iterations stay out of the chaotic regime (see the cfg1 tests)
    ref, _ = o.train(enwik6, iters, stride=stride, lr=lr)
    got = g.train_text(iters, stride=stride, lr=lr)
    nz = ref > 0
    assert np.max(np.abs(got[nz] - ref[nz]) / ref[nz]) < 1e-4
    xo, to = o.window(); xg, tg = g.window()
    assert np.array_equal(xo[1:], xg[1:]) and np.array_equal(to[1:], tg[1:])
    assert np.array_equal(g.positions(), o.positions())


def test_window_kernel_bit_exact(alice):
    """K10 vs the literal shift loop, through warm-up (-1 columns), many strides and text wrap-around."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 16, 6, 5
    text = alice[:200]
    pos = [6, 50, 199, 120, 7]
    o = orc.Oracle(M, N, S, B, "f32"); o.set_positions(pos)
    g = el.LSTM(M, N, S, B); g.init_params(1); g.load_text(text); g.set_positions(pos)
    for stride in [1, 1, 1, 2, 5, 1, 3, 5, 5, 4] * 12:
        o.advance(text, stride)
        g.train_text(1, stride=stride, lr=0.0, want_losses=False)
        xo, to = o.window(); xg, tg = g.window()
        assert np.array_equal(xo[1:], xg[1:]) and np.array_equal(to[1:], tg[1:])


def test_enwik5_known_answer_on_gpu(golden_dir):
    """test() on the reference's trained N=32 checkpoint: logged 3.24396 bits/char
    (OV/lstm_eigen_class_CUDA/models/enwik5_test.txt:1)."""
    import eigen_lstm_b200 as el
    z = np.load(os.path.join(golden_dir, "enwik5_test.npz"))
    params = [z["W"], z["U"], z["b"], z["Why"], z["by"]]
    g = el.LSTM(256, 32, 2, 1); g.set_params(params)
    bpc = g.test(z["test_bytes"].tobytes())
    assert abs(bpc - float(z["logged_test_bpc"])) < 2e-4
    o = orc.Oracle(256, 32, 2, 1, "f32"); o.set_params(params)
    assert abs(bpc - o.eval_bpc(z["test_bytes"].tobytes())) < 1e-5


def test_untrained_bpc_is_8_on_gpu(enwik6):
    import eigen_lstm_b200 as el
    g = el.LSTM(256, 32, 3, 4); g.init_params(1, 0.01, 1.0)
    assert abs(g.test(enwik6[:2000]) - 8.0) < 5e-3


def test_sampling_matches_oracle(golden_dir):
    """sample(): same uniforms (mt19937(seed)), same weights -> same bytes (R/lstm.cc:293-356)."""
    import eigen_lstm_b200 as el
    z = np.load(os.path.join(golden_dir, "enwik5_test.npz"))
    params = [z["W"], z["U"], z["b"], z["Why"], z["by"]]
    N = 32
    g = el.LSTM(256, N, 2, 1); g.set_params(params)
    o = orc.Oracle(256, N, 2, 1, "f32"); o.set_params(params)
    h0 = orc.randn(N, 1, 0, 0.1, 5).ravel(); c0 = orc.randn(N, 1, 0, 0.1, 6).ravel()
    n = 400
    assert np.array_equal(g.sample(n, seed=11, h0=h0, c0=c0, greedy=True), o.sample(h0, c0, 11, n, greedy=True))
    a = g.sample(n, seed=11, h0=h0, c0=c0); b = o.sample(h0, c0, 11, n)
    assert (a == b).mean() > 0.99, (a == b).mean()   # a draw within 1 ulp of a cdf edge may flip one byte


def test_init_params_matches_reference_randn():
    import eigen_lstm_b200 as el
    M, N = 256, 16
    g = el.LSTM(M, N, 3, 2); g.init_params(42, 0.01, 1.0)
    ref = orc.init_params(M, N, seed=42, sd=0.01, forget_bias=1.0)
    for a, b in zip(g.params(), ref):
        assert np.array_equal(a, b)


def test_text_checkpoint_roundtrip_and_format(tmp_path):
    """Parameters::save_to_disk / load_from_disk text format (OV/lstm_eigen_class_CUDA/io.h:16-81)."""
    import eigen_lstm_b200 as el
    M, N = 256, 8
    g = el.LSTM(M, N, 3, 1); g.init_params(3, 0.3, 1.0)
    prefix = str(tmp_path / "ck")
    g.save_to_disk(prefix)
    before = g.params()
    for name, p in zip(orc.NAMES, before):
        txt = open(f"{prefix}_{name}.txt").read()
        assert not txt.endswith("\n")
        rows = txt.split("\n")
        assert len(rows) == p.shape[0] and len(set(len(r) for r in rows)) == 1  # aligned columns
        parsed = np.loadtxt(f"{prefix}_{name}.txt", ndmin=2)
        assert parsed.shape == p.shape
        assert np.allclose(parsed, p, rtol=1e-5, atol=0)
    h = el.LSTM(M, N, 3, 1); h.load_from_disk(prefix)
    for a, b in zip(h.params(), before):
        assert np.allclose(a, b, rtol=1e-5, atol=0)
    with pytest.raises(el.LstmError):
        h.load_from_disk(str(tmp_path / "missing"))


def test_bin_checkpoint_resume_bit_exact(alice, tmp_path):
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 32, 4, 2
    a = el.LSTM(M, N, S, B); a.init_params(5); a.load_text(alice); a.set_positions([10, 3000])
    a.train_text(25, 1, 0.1)
    path = str(tmp_path / "ck.bin"); a.save_bin(path)
    la = a.train_text(15, 1, 0.1)
    b = el.LSTM(M, N, S, B); b.load_text(alice); b.load_bin(path)
    lb = b.train_text(15, 1, 0.1)
    assert np.array_equal(la, lb)


def test_errors_are_loud():
    import eigen_lstm_b200 as el
    g = el.LSTM(256, 8, 3, 1)
    with pytest.raises(el.LstmError):
        g.backward()                       # before forward
    with pytest.raises(el.LstmError):
        g.train_text(1)                    # no text
    with pytest.raises(el.LstmError):
        g.forward(np.full((3, 1), 300), np.zeros((3, 1)))   # byte out of range
    with pytest.raises(el.LstmError):
        el.LSTM(300, 8, 3, 1)              # M > 256


# ---- K9: persistent multi-CTA batch-1 recurrence (used for N >= 128) -------------------------------------------
@pytest.mark.parametrize("N", [128, 200, 256])
def test_persistent_eval_matches_oracle(enwik6, N):
    """test() on the cooperative persistent kernel (weights resident in shared memory, one grid barrier per
    character) vs the oracle's serial recipe."""
    import eigen_lstm_b200 as el
    params = orc.init_params(256, N, seed=31, sd=0.08, forget_bias=1.0)
    g = el.LSTM(256, N, 2, 1); g.set_params(params)
    o = orc.Oracle(256, N, 2, 1, "f32"); o.set_params(params)
    text = enwik6[1000:1700]
    assert abs(g.test(text) - o.eval_bpc(text)) < 2e-5


def test_persistent_eval_chains_state_across_chunks(enwik6):
    """40 000 bytes = three 16 384-character launches with h, c carried on the device."""
    import eigen_lstm_b200 as el
    N = 128
    params = orc.init_params(256, N, seed=32, sd=0.08)
    g = el.LSTM(256, N, 2, 1); g.set_params(params)
    o = orc.Oracle(256, N, 2, 1, "f32"); o.set_params(params)
    text = enwik6[:40000]
    assert abs(g.test(text) - o.eval_bpc(text)) < 2e-5


@pytest.mark.parametrize("N", [128, 200])
def test_persistent_sampling_matches_oracle(N):
    import eigen_lstm_b200 as el
    params = orc.init_params(256, N, seed=33, sd=0.12)
    g = el.LSTM(256, N, 2, 1); g.set_params(params)
    o = orc.Oracle(256, N, 2, 1, "f32"); o.set_params(params)
    h0 = orc.randn(N, 1, 0, 0.1, 5).ravel(); c0 = orc.randn(N, 1, 0, 0.1, 6).ravel()
    n = 400
    assert np.array_equal(g.sample(n, seed=3, h0=h0, c0=c0, greedy=True), o.sample(h0, c0, 3, n, greedy=True))
    a = g.sample(n, seed=3, h0=h0, c0=c0); b = o.sample(h0, c0, 3, n)
    assert (a == b).mean() > 0.99, (a == b).mean()
