"""CPU-side checks of the boundary: the library loads, exports every symbol include/lstm_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import os
import re

import pytest

from tests.conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "lstm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lstm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from eigen_lstm_b200 import _lib, build
    build.build()
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lstm_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, set(_lib.SYMBOLS) ^ set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import eigen_lstm_b200 as el
    with pytest.raises(el.LstmError, match="no CUDA device"):
        el.LSTM(256, 64, 3, 1)


def test_product_does_not_touch_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "eigen_lstm_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
