"""ctypes loader for liblstm_b200.so (the C ABI of include/lstm_b200.h).

Fails loudly if the CUDA library is missing: there is no Python / CPU fallback for the hot path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblstm_b200.so")

# every symbol include/lstm_b200.h declares: name -> (restype, argtypes)
_vp, _i, _sz, _u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
_fp, _dp, _ip, _bp = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
SYMBOLS = {
    "lstm_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i]),
    "lstm_destroy": (_i, [_vp]),
    "lstm_last_error": (C.c_char_p, [_vp]),
    "lstm_sync": (_i, [_vp]),
    "lstm_set_option": (_i, [_vp, _i, C.c_double]),
    "lstm_debug_variant": (_i, [_vp, _vp]),
    "lstm_tensor_size": (C.c_long, [_vp, _i]),
    "lstm_set_tensor": (_i, [_vp, _i, _i, _vp, _sz]),
    "lstm_get_tensor": (_i, [_vp, _i, _i, _vp, _sz]),
    "lstm_init_params": (_i, [_vp, _u64, C.c_float, C.c_float]),
    "lstm_set_state": (_i, [_vp, _vp, _vp]),
    "lstm_get_state": (_i, [_vp, _vp, _vp]),
    "lstm_reset_state": (_i, [_vp, _u64, C.c_float]),
    "lstm_forward": (_i, [_vp, _vp, _vp, _dp]),
    "lstm_backward": (_i, [_vp]),
    "lstm_adagrad": (_i, [_vp, C.c_float, C.c_double, C.c_float]),
    "lstm_carry_state": (_i, [_vp, _i]),
    "lstm_train_step": (_i, [_vp, _vp, _vp, _i, C.c_float, _dp]),
    "lstm_load_text": (_i, [_vp, _vp, _sz]),
    "lstm_set_positions": (_i, [_vp, _vp]),
    "lstm_get_positions": (_i, [_vp, _vp]),
    "lstm_train_text": (_i, [_vp, _i, _i, C.c_float, _vp]),
    "lstm_get_window": (_i, [_vp, _vp, _vp]),
    "lstm_eval_bpc": (_i, [_vp, _vp, _sz, _dp]),
    "lstm_sample": (_i, [_vp, _u64, _vp, _vp, _vp, _sz, _i]),
    "lstm_get_activation": (_i, [_vp, _i, _i, _vp, _sz]),
    "lstm_gradcheck": (_i, [_vp, _vp, _vp, _i, _u64, C.c_double, _vp, _vp, _vp, _vp, C.POINTER(_i)]),
    "lstm_save_text_ckpt": (_i, [_vp, C.c_char_p]),
    "lstm_load_text_ckpt": (_i, [_vp, C.c_char_p]),
    "lstm_save_bin": (_i, [_vp, C.c_char_p]),
    "lstm_load_bin": (_i, [_vp, C.c_char_p]),
    "lstm_dp_unique_id": (_i, [_vp]),
    "lstm_dp_init": (_i, [_vp, _i, _i, _vp]),
    "lstm_debug_kernel_clocks": (_i, [_vp, _vp]),
    "lstm_set_profiling": (_i, [_vp, _i]),
    "lstm_get_phase_ms": (_i, [_vp, _vp]),
    "lstm_launch_count": (C.c_long, [_vp]),
    "lstm_stream": (_vp, [_vp]),
    "lstm_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the shared library (building is the job of eigen_lstm_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m eigen_lstm_b200.build` "
            "(the CUDA library is the product; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
