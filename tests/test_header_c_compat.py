"""include/lstm_b200.h is a plain C header: a C99 translation unit that includes it and references every entry point
must compile (no C++ or torch types across the boundary)."""
import os
import re
import subprocess

from tests.conftest import ROOT


def test_header_compiles_as_c99(tmp_path):
    hdr = open(os.path.join(ROOT, "include", "lstm_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(lstm_[a-z0-9_]+)\s*\(", body)))
    src = tmp_path / "t.c"
    src.write_text('#include "lstm_b200.h"\ntypedef void (*fn_t)(void);\nfn_t table[] = {' + ", ".join(f"(fn_t){n}" for n in names) + "};\nint main(void) { return 0; }\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
