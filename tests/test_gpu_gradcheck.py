"""The reference's signature self-test, made automatic: analytic gradients of the CUDA path against central
differences of a DOUBLE-precision forward (sampled ~100 entries per tensor, delta = 1e-5), with the reference's own
acceptance rule (OV/lstm_eigen_class_batch/lstm.h:203-261, lstm.cc:440-510: rel.err = |a-n|/|a+n|, fail if
max > 1e-1 or mean > 1e-3).  The double forward is the oracle's; the analytic gradients are the GPU's (fp32 path)."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
LN2 = np.log(2.0)


def test_gpu_analytic_gradients_pass_the_reference_gradient_check():
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 20, 6, 4        # OV/lstm_eigen_class defaults: N = 20
    rng = np.random.default_rng(0)
    params = [rng.normal(0, 0.2, s).astype(np.float32) for s in [(4 * N, M), (4 * N, N), (4 * N, 1), (M, N), (M, 1)]]
    x = rng.integers(0, M, (S, B)).astype(np.int32)
    t = rng.integers(0, M, (S, B)).astype(np.int32)
    h0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    c0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    g = el.LSTM(M, N, S, B)
    g.set_params(params); g.set_state(h0, c0)
    g.forward(x, t); g.backward()
    ana = g.grads()
    o = orc.Oracle(M, N, S, B, "f64")
    o.set_params(params); o.set_window(x, t); o.set_state("h", 0, h0); o.set_state("c", 0, c0)
    delta = 1e-5
    worst, errs = 0.0, []
    for w in range(5):
        p = params[w].astype(np.float64)
        if w == 0:   # W: only the columns whose byte occurs in the window have a non-zero gradient; sample those
            cols = np.unique(x[1:][x[1:] >= 0])
            idxs = [(int(rng.integers(0, 4 * N)), int(rng.choice(cols))) for _ in range(100)]
        else:
            flat = rng.choice(p.size, size=min(100, p.size), replace=False)
            idxs = [np.unravel_index(int(f), p.shape) for f in flat]
        for idx in idxs:
            q = p.copy(); q[idx] += delta; o.set(orc.PARAM, w, q); lp = o.forward()
            q = p.copy(); q[idx] -= delta; o.set(orc.PARAM, w, q); lm = o.forward()
            n = (lp - lm) * LN2 * B / (2 * delta)      # forward() is bits / B; the gradients are those of the summed ln-loss
            a = float(ana[w][idx])
            if abs(a) + abs(n) < 1e-6:
                continue
            r = abs(a - n) / abs(a + n)
            errs.append(r); worst = max(worst, r)
        o.set(orc.PARAM, w, p)
    assert len(errs) > 300
    assert worst < 1e-1 and float(np.mean(errs)) < 1e-3, (worst, float(np.mean(errs)))     # the reference's rule
    assert float(np.mean(errs)) < 1e-4                                                      # what fp32 actually holds


def test_device_gradcheck_api_agrees_with_the_double_oracle_and_passes():
    """lstm_gradcheck (the product's own gradient check: fp64 perturbed forward passes on the device) must (a) produce the
    same central differences as the double-precision oracle at the same entries and (b) pass the reference's rule on the
    fp32 path."""
    import eigen_lstm_b200 as el
    M, N, S, B = 256, 20, 6, 4
    rng = np.random.default_rng(1)
    params = [rng.normal(0, 0.2, s).astype(np.float32) for s in [(4 * N, M), (4 * N, N), (4 * N, 1), (M, N), (M, 1)]]
    x = rng.integers(0, M, (S, B)).astype(np.int32)
    t = rng.integers(0, M, (S, B)).astype(np.int32)
    x[1, 0] = -1                                  # a warm-up (all-zero) INPUT column is a legitimate part of the function
    h0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    c0 = rng.normal(0, 0.1, (N, B)).astype(np.float32)
    g = el.LSTM(M, N, S, B)
    g.set_params(params); g.set_state(h0, c0)
    # an all-zero TARGET column is refused: there the reference's dy = probs - 0 is not the gradient of its loss
    bad = t.copy(); bad[2, 1] = -1
    with pytest.raises(el.LstmError):
        g.gradcheck(x, bad, per_tensor=4)
    per = 40
    res = g.gradcheck(x, t, per_tensor=per, seed=3, delta=1e-5)
    assert res["passed"], res["report"]
    for name, r in res["report"].items():
        assert r["max"] < 1e-1 and r["mean"] < 1e-3, (name, r)
    # the same central differences from the oracle's double forward
    o = orc.Oracle(M, N, S, B, "f64")
    o.set_params(params); o.set_window(x, t); o.set_state("h", 0, h0); o.set_state("c", 0, c0)
    delta = 1e-5
    checked = 0
    for w in range(5):
        p = params[w].astype(np.float64)
        for q in range(per):
            i = int(res["idx"][w, q])
            if i < 0:
                continue
            idx = np.unravel_index(i, p.shape, order="F")
            qq = p.copy(); qq[idx] += delta; o.set(orc.PARAM, w, qq); lp = o.forward()
            qq = p.copy(); qq[idx] -= delta; o.set(orc.PARAM, w, qq); lm = o.forward()
            n_ref = (lp - lm) * LN2 * B / (2 * delta)
            n_dev = float(res["numeric"][w, q])
            assert abs(n_dev - n_ref) <= 1e-6 * max(1.0, abs(n_ref)), (w, i, n_dev, n_ref)
            checked += 1
        o.set(orc.PARAM, w, p)
    assert checked >= 4 * per


def test_lstm_binary_gradcheck_mode(tmp_path, alice):
    """`lstm --gradcheck` prints the reference's per-tensor report (check_gradient_error) and exits 0 when it passes."""
    import os
    import subprocess
    from tests.conftest import ROOT
    (tmp_path / "alice29.txt").write_bytes(alice[:3000])
    out = subprocess.run([os.path.join(ROOT, "eigen_lstm_b200", "lstm"), "--gradcheck", "--seed", "5", "--hidden", "32", "--seq", "8",
                          "--batch", "4"], cwd=tmp_path, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
    so = out.stdout.decode()
    assert out.returncode == 0, so + out.stderr.decode()
    for name in ["U", "W", "Why", "b", "by"]:
        assert f"[{name}]" in so
    assert so.count("max rel. error") == 5 and so.count("mean rel. error") == 5 and "gradient check OK" in so
