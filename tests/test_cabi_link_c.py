"""A plain-C program links against liblstm_b200.so through include/lstm_b200.h and, on a box without a GPU, gets the
documented loud failure from lstm_create (no CPU fallback) instead of a crash."""
import os
import subprocess

from tests.conftest import ROOT

SRC = r'''
#include <stdio.h>
#include <string.h>
#include "lstm_b200.h"
int main(void) {
  lstm_ctx* ctx = NULL;
  int rc = lstm_create(&ctx, 256, 64, 3, 1, 0, LSTM_F32);
  printf("version=%s rc=%d\n", lstm_version(), rc);
  if (rc != 0) { printf("error=%s\n", lstm_last_error(NULL)); return ctx == NULL ? 0 : 3; }
  { unsigned char text[64]; double losses[4]; memset(text, 'a', sizeof text);
    if (lstm_init_params(ctx, 1, 0.01f, 0.0f) || lstm_load_text(ctx, text, sizeof text) || lstm_train_text(ctx, 4, 1, 0.1f, losses)) return 4;
    printf("loss0=%f launches=%ld\n", losses[0], lstm_launch_count(ctx)); }
  return lstm_destroy(ctx);
}
'''


def test_c_program_links_and_fails_loudly_without_gpu(tmp_path):
    from eigen_lstm_b200 import build
    build.build()
    libdir = os.path.join(ROOT, "eigen_lstm_b200")
    src = tmp_path / "t.c"
    src.write_text(SRC)
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + libdir, "-llstm_b200", "-Wl,-rpath," + libdir])
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr)
    assert "version=eigen_lstm_b200" in r.stdout
    import torch
    if not torch.cuda.is_available():
        assert "rc=-2" in r.stdout and "no CUDA device" in r.stdout
    else:
        assert "rc=0" in r.stdout and "loss0=" in r.stdout
