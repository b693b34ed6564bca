"""Edge cases of the boundary: the shortest window (S = 2: one active timestep), one stream, windows made only of
not-yet-filled columns, maximum byte values, and argument errors."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import make_pair, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,tol", [(0, 2e-5), (1, 6e-2)])
def test_shortest_window_single_stream(dtype, tol):
    M, N, S, B = 256, 64, 2, 1
    o, g, _ = make_pair(M, N, S, B, seed=8, sd=0.1, dtype=dtype)
    x = np.array([[-1], [255]], np.int32)          # largest byte value
    t = np.array([[-1], [0]], np.int32)
    o.set_window(x, t)
    lo, lg = o.forward(), g.forward(x, t)
    assert abs(lg - lo) <= (2e-6 if dtype == 0 else 2e-2) * abs(lo)
    o.backward(); g.backward()
    for name, a, b in zip(orc.NAMES, g.grads(), o.grads()):
        err = rel_err(a, b) if dtype == 0 else float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        assert err < tol, name


@pytest.mark.parametrize("dtype", [0, 1])
def test_all_null_window_gives_zero_loss_and_dy_equal_probs(dtype):
    """The reference's very first iterations have all-zero input columns and all-zero target columns
    (R/lstm.cc:84,124,169-170): loss contribution 0, but dy = probs still drives the gradients (R/lstm.cc:225)."""
    M, N, S, B = 256, 64, 4, 2
    o, g, _ = make_pair(M, N, S, B, seed=9, sd=0.1, dtype=dtype)
    x = np.full((S, B), -1, np.int32); t = np.full((S, B), -1, np.int32)
    o.set_window(x, t)
    assert o.forward() == 0.0 and g.forward(x, t) == 0.0
    o.backward(); g.backward()
    gw = g.grads()
    assert np.all(gw[0] == 0)                      # no input byte anywhere: dW = 0 exactly
    for name, a, b in zip(orc.NAMES[1:], gw[1:], o.grads()[1:]):
        err = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        assert err < (2e-5 if dtype == 0 else 6e-2), name


def test_argument_errors():
    import eigen_lstm_b200 as el
    g = el.LSTM(256, 16, 5, 2)
    with pytest.raises(el.LstmError):
        g.load_text(b"abcd")                       # shorter than the window
    g.load_text(bytes(range(256)) * 2)
    with pytest.raises(el.LstmError):
        g.set_positions([1, 10])                   # position < S
    with pytest.raises(el.LstmError):
        g.set_positions([5, 512])                  # position >= length
    with pytest.raises(el.LstmError):
        g.activation("g", 0)                       # gates exist for t >= 1 only
    with pytest.raises(el.LstmError):
        el.LSTM(256, 16, 1, 1)                     # S < 2
    with pytest.raises(el.LstmError):
        el.LSTM(256, 16, 3, 1, device=99)
