"""eigen_lstm_b200 — B200-native (sm_100a) character-LSTM training path behind a C ABI.

The product is eigen_lstm_b200/liblstm_b200.so (include/lstm_b200.h).  This package holds the CUDA
sources (csrc/), the in-tree build (build.py) and a thin ctypes mirror of the reference's class layer
(model.py).  Importing the package does not load CUDA; creating an `LSTM` does, and fails loudly
without the library or a GPU.
"""
from .model import BF16, F32, GRAD, MEM, PARAM, LSTM, LstmError, dp_unique_id  # noqa: F401
