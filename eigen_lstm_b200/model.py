"""Host-side mirror of the reference's class layer, on top of the C ABI.

Names follow the reference's last snapshots so the parity tests read like its own checks:

    Parameters{W,U,b,Why,by}, reset(), save_to_disk/load_from_disk   OV/lstm_eigen_class_CUDA/lstm.h:43-112
    LSTM<S>: forward(), backward(), state arrays h,c,g,probs         OV/lstm_eigen_class_CUDA/lstm.h:114-397
    adagrad(p, d, m, lr)                                             OV/lstm_eigen_class_CUDA/lstm.cc:397-417
    test(p, data), sample(p, n)                                      OV/lstm_eigen_class_CUDA/lstm.cc:578-720

All arithmetic happens in liblstm_b200.so on the GPU; numpy is only the container for host copies
(column-major, mathematical shape (rows, cols)).
"""
import ctypes as C

import numpy as np

from . import _lib

W, U, B_, WHY, BY = 0, 1, 2, 3, 4
PARAM, GRAD, MEM = 0, 1, 2
F32, BF16 = 0, 1
NAMES = ["W", "U", "b", "Why", "by"]
_ACT = {"h": 0, "c": 1, "g": 2, "probs": 3, "dhy": 4, "dg": 5}


class LstmError(RuntimeError):
    pass


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class LSTM:
    """One context = Parameters + LSTM<S> state + Adagrad memory on one GPU."""

    def __init__(self, M, N, S, B, device=0, dtype=F32):
        self.lib = _lib.load()
        self.M, self.N, self.S, self.B = M, N, S, B
        self.ctx = C.c_void_p()
        rc = self.lib.lstm_create(C.byref(self.ctx), M, N, S, B, device, dtype)
        if rc != 0:
            raise LstmError(f"lstm_create failed ({rc}): {self.lib.lstm_last_error(None).decode()}")

    def close(self):
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.lstm_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise LstmError(f"error {rc}: {self.lib.lstm_last_error(self.ctx).decode()}")

    # ---- Parameters ----
    def shape(self, which):
        M, N = self.M, self.N
        return [(4 * N, M), (4 * N, N), (4 * N, 1), (M, N), (M, 1)][which]

    def set(self, kind, which, arr):
        a = np.asfortranarray(np.asarray(arr, dtype=np.float32).reshape(self.shape(which)))
        self._ck(self.lib.lstm_set_tensor(self.ctx, kind, which, _ptr(a), a.size))

    def get(self, kind, which):
        a = np.empty(self.shape(which), dtype=np.float32, order="F")
        self._ck(self.lib.lstm_get_tensor(self.ctx, kind, which, _ptr(a), a.size))
        return a

    def set_params(self, params):
        for i, p in enumerate(params):
            self.set(PARAM, i, p)

    def params(self):
        return [self.get(PARAM, i) for i in range(5)]

    def grads(self):
        return [self.get(GRAD, i) for i in range(5)]

    def adagrad_mem(self):
        return [self.get(MEM, i) for i in range(5)]

    def init_params(self, seed, std=0.01, forget_bias=0.0):
        self._ck(self.lib.lstm_init_params(self.ctx, seed, std, forget_bias))

    def save_to_disk(self, prefix):
        self._ck(self.lib.lstm_save_text_ckpt(self.ctx, prefix.encode()))

    def load_from_disk(self, prefix):
        self._ck(self.lib.lstm_load_text_ckpt(self.ctx, prefix.encode()))

    def save_bin(self, path):
        self._ck(self.lib.lstm_save_bin(self.ctx, path.encode()))

    def load_bin(self, path):
        self._ck(self.lib.lstm_load_bin(self.ctx, path.encode()))

    # ---- state ----
    def set_state(self, h0, c0):
        h = np.asfortranarray(np.asarray(h0, dtype=np.float32).reshape(self.N, self.B))
        c = np.asfortranarray(np.asarray(c0, dtype=np.float32).reshape(self.N, self.B))
        self._ck(self.lib.lstm_set_state(self.ctx, _ptr(h), _ptr(c)))

    def get_state(self):
        h = np.empty((self.N, self.B), dtype=np.float32, order="F")
        c = np.empty((self.N, self.B), dtype=np.float32, order="F")
        self._ck(self.lib.lstm_get_state(self.ctx, _ptr(h), _ptr(c)))
        return h, c

    def reset_state(self, seed=0, std=0.0):
        self._ck(self.lib.lstm_reset_state(self.ctx, seed, std))

    def activation(self, what, t):
        rows = {"h": self.N, "c": self.N, "g": 4 * self.N, "probs": self.M, "dhy": self.N, "dg": 4 * self.N}[what]
        a = np.empty((rows, self.B), dtype=np.float32, order="F")
        self._ck(self.lib.lstm_get_activation(self.ctx, _ACT[what], t, _ptr(a), a.size))
        return a

    # ---- LSTM<S>::forward / backward / adagrad ----
    def _win(self, a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(self.S, self.B))

    def forward(self, x_idx, t_idx):
        x, t = self._win(x_idx), self._win(t_idx)
        loss = C.c_double()
        self._ck(self.lib.lstm_forward(self.ctx, _ptr(x), _ptr(t), C.byref(loss)))
        return loss.value

    def backward(self):
        self._ck(self.lib.lstm_backward(self.ctx))

    def gradcheck(self, x_idx, t_idx, per_tensor=100, seed=0, delta=1e-5):
        """compute_all_numerical_grads + check_gradients (OV/lstm_eigen_class_batch/lstm.h:203-261, lstm.cc:440-510):
        analytic gradients of this context vs double-precision central differences evaluated on the device.
        Returns dict(passed, report{name: dict(max, mean, n_range, a_range)}, idx, numeric, analytic) with the three
        arrays shaped (5, per_tensor) (idx = column-major index inside the tensor, -1 = unused slot)."""
        x, t = self._win(x_idx), self._win(t_idx)
        rep = np.zeros((5, 6), dtype=np.float64)
        idx = np.full((5, per_tensor), -1, dtype=np.int64)
        num = np.zeros((5, per_tensor), dtype=np.float64)
        ana = np.zeros((5, per_tensor), dtype=np.float64)
        ok = C.c_int(0)
        self._ck(self.lib.lstm_gradcheck(self.ctx, _ptr(x), _ptr(t), int(per_tensor), int(seed), float(delta), _ptr(rep),
                                         _ptr(idx), _ptr(num), _ptr(ana), C.byref(ok)))
        report = {NAMES[w]: dict(max=rep[w, 0], mean=rep[w, 1], n_range=(rep[w, 2], rep[w, 3]), a_range=(rep[w, 4], rep[w, 5]))
                  for w in range(5)}
        return dict(passed=bool(ok.value), report=report, idx=idx, numeric=num, analytic=ana)

    def adagrad(self, lr=0.1, eps=1e-10, clip=0.0):
        self._ck(self.lib.lstm_adagrad(self.ctx, lr, eps, clip))

    def carry(self, stride=1):
        self._ck(self.lib.lstm_carry_state(self.ctx, stride))

    def train_step(self, x_idx, t_idx, stride=1, lr=0.1, want_loss=True):
        x, t = self._win(x_idx), self._win(t_idx)
        loss = C.c_double()
        self._ck(self.lib.lstm_train_step(self.ctx, _ptr(x), _ptr(t), stride, lr, C.byref(loss) if want_loss else None))
        if want_loss:
            self.sync()                  # the step is asynchronous: its loss arrives with the next synchronisation
        return loss.value if want_loss else None

    def train_step_async(self, x_idx, t_idx, losses, i, stride=1, lr=0.1):
        """lstm_train_step without waiting: losses (float64 numpy array that outlives the next sync()) gets the step's loss
        in losses[i] when sync() is called."""
        x, t = self._win(x_idx), self._win(t_idx)
        assert losses.dtype == np.float64 and losses.flags["C_CONTIGUOUS"]
        slot = C.cast(C.c_void_p(losses.ctypes.data + 8 * i), C.POINTER(C.c_double))
        self._ck(self.lib.lstm_train_step(self.ctx, _ptr(x), _ptr(t), stride, lr, slot))

    # ---- device text pipeline ----
    def load_text(self, data):
        d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        self._ck(self.lib.lstm_load_text(self.ctx, _ptr(d), d.size))

    def set_positions(self, pos):
        p = np.ascontiguousarray(np.asarray(pos, dtype=np.uint64))
        assert p.size == self.B
        self._ck(self.lib.lstm_set_positions(self.ctx, _ptr(p)))

    def positions(self):
        p = np.empty(self.B, dtype=np.uint64)
        self._ck(self.lib.lstm_get_positions(self.ctx, _ptr(p)))
        return p

    def window(self):
        x = np.empty((self.S, self.B), dtype=np.int32)
        t = np.empty((self.S, self.B), dtype=np.int32)
        self._ck(self.lib.lstm_get_window(self.ctx, _ptr(x), _ptr(t)))
        return x, t

    def train_text(self, iters, stride=1, lr=0.1, want_losses=True):
        losses = np.zeros(iters, dtype=np.float64) if want_losses else None
        self._ck(self.lib.lstm_train_text(self.ctx, iters, stride, lr, _ptr(losses) if want_losses else None))
        return losses

    # ---- test() / sample() ----
    def test(self, data):
        d = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8))
        out = C.c_double()
        self._ck(self.lib.lstm_eval_bpc(self.ctx, _ptr(d), d.size, C.byref(out)))
        return out.value

    def sample(self, n, seed=0, h0=None, c0=None, greedy=False):
        out = np.zeros(n, dtype=np.uint8)
        h = None if h0 is None else np.ascontiguousarray(np.asarray(h0, dtype=np.float32).reshape(self.N))
        c = None if c0 is None else np.ascontiguousarray(np.asarray(c0, dtype=np.float32).reshape(self.N))
        self._ck(self.lib.lstm_sample(self.ctx, seed, None if h is None else _ptr(h), None if c is None else _ptr(c),
                                      _ptr(out), n, int(bool(greedy))))
        return out

    # ---- data parallel / measurement ----
    def dp_init(self, rank, world, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        self._ck(self.lib.lstm_dp_init(self.ctx, rank, world, buf))

    def sync(self):
        self._ck(self.lib.lstm_sync(self.ctx))

    def set_profiling(self, on=True):
        self._ck(self.lib.lstm_set_profiling(self.ctx, int(on)))

    def phase_ms(self):
        a = np.zeros(16, dtype=np.float32)
        self._ck(self.lib.lstm_get_phase_ms(self.ctx, _ptr(a)))
        names = ["window", "fwd_recurrence", "logits_softmax", "dhy_dwhy_gemms", "bwd_recurrence", "weight_grads",
                 "allreduce_wait", "adagrad", "total"]
        out = dict(zip(names, a[:9].tolist()))
        if a[9:].any():   # data parallel: ms since the start of the iteration
            out["comm_timeline"] = {"summed_at": dict(zip(["tail_panel", "why_by", "panel0", "panel1"], a[9:13].tolist())),
                                    "handed_over_at": dict(zip(["panel0", "panel1", "tail_panel"], a[13:16].tolist()))}
        return out

    # ---- options of the training path (include/lstm_b200.h LSTM_OPT_*) ----
    def set_clip(self, clip):
        self._ck(self.lib.lstm_set_option(self.ctx, 0, float(clip)))

    def set_loss_mode(self, mode):
        """'log2-all' (R/lstm.cc:204-207) or 'last-ln' (OV/lstm_eigen_class_batch/lstm.cc:308-319)."""
        self._ck(self.lib.lstm_set_option(self.ctx, 1, float({"log2-all": 0, "last-ln": 1}[mode])))

    def set_softmax_shift(self, mode):
        """'none' (R/lstm.cc:199-201) or 'global' (OV/lstm_eigen_class_batch/lstm.h:175)."""
        self._ck(self.lib.lstm_set_option(self.ctx, 2, float({"none": 0, "global": 1}[mode])))

    def variant(self):
        """Kernel instantiations this context's shape selects, see lstm_debug_variant."""
        a = np.zeros(8, dtype=np.int32)
        self._ck(self.lib.lstm_debug_variant(self.ctx, _ptr(a)))
        keys = ["fwd_bn", "fwd_pair", "bwd_bn", "bwd_variant", "wgrad_bn", "fwd_persistent", "bwd_persistent", "train_small"]
        return dict(zip(keys, a.tolist()))

    def launch_count(self):
        return int(self.lib.lstm_launch_count(self.ctx))

    def stream(self):
        return self.lib.lstm_stream(self.ctx)


def dp_unique_id():
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    rc = lib.lstm_dp_unique_id(buf)
    if rc != 0:
        raise LstmError(f"lstm_dp_unique_id failed ({rc})")
    return bytes(buf)
