"""Golden OUTPUT of the reference's own source: run oracle/_ref/lstm_ref and record what it prints.

    make -C oracle ref && python tests/golden/make_ref_run.py        (in the build container)

oracle/_ref/lstm_ref is the UNMODIFIED /root/reference/lstm.cc compiled against oracle/eigen_shim (the real Eigen
is not installed; see oracle/Makefile) with std::random_device replaced by a seed counter (REF_SEED).  The program
has no arguments: it reads "alice29.txt" from its working directory and runs 1000 epochs, so it is started in a
scratch directory holding the first CORPUS_BYTES bytes of R/alice29.txt (= tests/golden/alice29_head.bin[:3000])
and stopped after EPOCHS epoch reports.

The batched snapshot OV/lstm_eigen_BLAS/lstm.cc (B = 4 streams, S = 3; built without -DUSE_BLAS = its pure-Eigen
branch) is recorded the same way as oracle/_ref/lstm_blas_ref on the first 2000 bytes of R/enwik5.txt
(= tests/golden/enwik6_head.bin[:2000], the corpora share their prefix) -> tests/golden/ref_lstm_eigen_blas_run.json.
Its stream start positions come from the C library's unseeded rand() (OV/lstm_eigen_BLAS/lstm.cc:150-154); the
values it drew are stored in the fixture (`positions`) so the replay does not depend on the libc.

The double-precision class snapshot OV/lstm_eigen_class_batch/lstm.cc (+ lstm.h; N = 32, S = 5, B = 4, 95 % of
alice29.txt for training) likewise as oracle/_ref/lstm_class_batch_ref -> tests/golden/ref_lstm_class_batch_run.json;
besides the epoch lines and samples it records the running loss the program prints every 100 iterations with 6
decimals (`progress_loss`) and the stream positions of every epoch (again glibc rand(), continuing sequence).

Output: tests/golden/ref_lstm_cc_run.json
  seed, corpus_bytes, epochs
  read_line        the "Read <n> bytes (alice29.txt)" line                      (R/lstm.cc:398)
  epoch_lines      the printed "Epoch e/1000, t = ..., avg loss = ... bits/char" lines   (:284-291)
  avg_loss         the avg-loss field of each line (3 decimals, as printed)
  generated_b64    the 1000 sampled bytes of each epoch, base64                 (:293-356)
  progress_fields  the "%7.2f%%" progress fields printed during epoch 1         (:274-279)
"""
import base64
import json
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(ROOT, "oracle", "_ref", "lstm_ref")
BIN_BLAS = os.path.join(ROOT, "oracle", "_ref", "lstm_blas_ref")
SEED, CORPUS_BYTES, EPOCHS = 1234, 3000, 4
BLAS_SEED, BLAS_CORPUS_BYTES, BLAS_EPOCHS, BLAS_B, BLAS_S = 99, 2000, 3, 4, 3
END = b"| Generated text END ************"
BIN_CLASS = os.path.join(ROOT, "oracle", "_ref", "lstm_class_batch_ref")
CLASS_SEED, CLASS_CORPUS_BYTES, CLASS_EPOCHS, CLASS_B, CLASS_S = 42, 3000, 2, 4, 5
PROGRAMS = {   # binary, file name the program opens, committed corpus it is a prefix of
    "lstm.cc": (BIN, "alice29.txt", "alice29_head.bin"),
    "lstm.cc+fma": (os.path.join(ROOT, "oracle", "_ref", "lstm_ref_fma"), "alice29.txt", "alice29_head.bin"),
    "lstm_eigen_class_batch": (BIN_CLASS, "alice29.txt", "alice29_head.bin"),
    # lstm_segment.cc trains on the file called "lstm.h" in its working directory (:50): it gets prose, not source
    "lstm_segment": (os.path.join(ROOT, "oracle", "_ref", "lstm_segment_ref"), "lstm.h", "alice29_head.bin"),
    "lstm_eigen_opt": (os.path.join(ROOT, "oracle", "_ref", "lstm_eigen_opt_ref"), "alice29.txt", "alice29_head.bin"),
    "lstm_eigen_class": (os.path.join(ROOT, "oracle", "_ref", "lstm_eigen_class_ref"), "alice29.txt", "alice29_head.bin"),
    "lstm_eigen_BLAS": (BIN_BLAS, "enwik5.txt", "enwik6_head.bin"),
}


def run_reference(seed=SEED, corpus_bytes=CORPUS_BYTES, epochs=EPOCHS, program="lstm.cc"):
    """Start the reference program on the truncated corpus, read its stdout until `epochs` samples were printed, stop it."""
    binary, fname, corpus = PROGRAMS[program]
    text = open(os.path.join(HERE, corpus), "rb").read()[:corpus_bytes]
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, fname), "wb").write(text)
        p = subprocess.Popen([binary], cwd=d, stdout=subprocess.PIPE, env=dict(os.environ, REF_SEED=str(seed)))
        out = b""
        try:
            while out.count(END) < epochs:
                chunk = p.stdout.read1(65536)
                if not chunk:
                    break
                out += chunk
        finally:
            p.kill()
            p.wait()
    return out[: out.rfind(END) + len(END)] if out.count(END) > epochs else out


def parse(out):
    read_line = out.split(b"\n", 1)[0].decode()
    epoch_lines = [m.decode() for m in re.findall(rb"Epoch \d+/\d+, t = [^\n]*", out)]
    avg = [re.search(r"avg loss = ([0-9.]+) bits/char", l).group(1) for l in epoch_lines]
    gen = re.findall(rb"\*{12} Generated text \|(.*?)\| Generated text END \*{12}", out, flags=re.S)
    first_epoch = out[: out.find(b"====")]
    progress = [m.decode() for m in re.findall(rb"([ 0-9.]{7})%\r", first_epoch)]
    return read_line, epoch_lines, avg, gen, progress


def main():
    out = run_reference()
    read_line, epoch_lines, avg, gen, progress = parse(out)
    assert len(gen) >= EPOCHS and all(len(g) == 1000 for g in gen[:EPOCHS]), [len(g) for g in gen]
    doc = dict(seed=SEED, corpus_bytes=CORPUS_BYTES, epochs=EPOCHS, read_line=read_line, epoch_lines=epoch_lines[:EPOCHS],
               avg_loss=avg[:EPOCHS], generated_b64=[base64.b64encode(g).decode() for g in gen[:EPOCHS]],
               progress_fields=progress,
               how="oracle/_ref/lstm_ref = unmodified /root/reference/lstm.cc + oracle/eigen_shim, REF_SEED=%d, "
                   "cwd holding alice29.txt = first %d bytes of R/alice29.txt" % (SEED, CORPUS_BYTES))
    json.dump(doc, open(os.path.join(HERE, "ref_lstm_cc_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:EPOCHS]))

    # batched snapshot
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(1)                                     # the program never calls srand: rand() starts from seed 1
    positions = [libc.rand() % (BLAS_CORPUS_BYTES - BLAS_S) + BLAS_S for _ in range(BLAS_B)]
    out = run_reference(BLAS_SEED, BLAS_CORPUS_BYTES, BLAS_EPOCHS, program="lstm_eigen_BLAS")
    read_line, epoch_lines, avg, gen, _ = parse(out)
    assert len(gen) >= BLAS_EPOCHS and all(len(g) == 2000 for g in gen[:BLAS_EPOCHS]), [len(g) for g in gen]
    doc = dict(seed=BLAS_SEED, corpus_bytes=BLAS_CORPUS_BYTES, epochs=BLAS_EPOCHS, B=BLAS_B, S=BLAS_S, N=64, positions=positions,
               read_line=read_line, epoch_lines=epoch_lines[:BLAS_EPOCHS], avg_loss=avg[:BLAS_EPOCHS],
               generated_b64=[base64.b64encode(g).decode() for g in gen[:BLAS_EPOCHS]],
               how="oracle/_ref/lstm_blas_ref = unmodified OV/lstm_eigen_BLAS/lstm.cc (no -DUSE_BLAS) + oracle/eigen_shim, "
                   "REF_SEED=%d, cwd holding enwik5.txt = first %d bytes of R/enwik5.txt; positions = glibc rand() "
                   "from its default seed" % (BLAS_SEED, BLAS_CORPUS_BYTES))
    json.dump(doc, open(os.path.join(HERE, "ref_lstm_eigen_blas_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:BLAS_EPOCHS]))

    # double-precision class snapshot
    out = run_reference(CLASS_SEED, CLASS_CORPUS_BYTES, CLASS_EPOCHS, program="lstm_eigen_class_batch")
    read_line, epoch_lines, avg, gen, _ = parse(out)
    assert len(gen) >= CLASS_EPOCHS and all(len(g) == 1500 for g in gen[:CLASS_EPOCHS]), [len(g) for g in gen]
    train_len = 95 * (CLASS_CORPUS_BYTES // 100)
    libc.srand(1)
    positions = [[libc.rand() % (train_len - CLASS_S) + CLASS_S for _ in range(CLASS_B)] for _ in range(CLASS_EPOCHS)]
    per_epoch = len([i for i in range(CLASS_S, train_len) if (i + 1) % 100 == 0])
    prog = re.findall(rb"\[Epoch (\d+)/1000\]\s+[0-9.]+%.*?loss = ([0-9.]+)", out)
    progress = [p[1].decode() for p in prog if int(p[0]) <= CLASS_EPOCHS]
    assert len(progress) == per_epoch * CLASS_EPOCHS, (len(progress), per_epoch)
    doc = dict(seed=CLASS_SEED, corpus_bytes=CLASS_CORPUS_BYTES, train_bytes=train_len, epochs=CLASS_EPOCHS, B=CLASS_B, S=CLASS_S, N=32,
               positions=positions, read_line=read_line, split_line=out.split(b"\n")[1].decode(),
               epoch_lines=epoch_lines[:CLASS_EPOCHS], avg_loss=avg[:CLASS_EPOCHS], progress_loss=progress,
               generated_b64=[base64.b64encode(g).decode() for g in gen[:CLASS_EPOCHS]],
               how="oracle/_ref/lstm_class_batch_ref = unmodified OV/lstm_eigen_class_batch/lstm.cc + lstm.h (no -DUSE_BLAS) + "
                   "oracle/eigen_shim, REF_SEED=%d, cwd holding alice29.txt = first %d bytes of R/alice29.txt"
                   % (CLASS_SEED, CLASS_CORPUS_BYTES))
    json.dump(doc, open(os.path.join(HERE, "ref_lstm_class_batch_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:CLASS_EPOCHS]))

    # strided-window variant.  The corpus length is 3001 so that the last window of an epoch satisfies the program's
    # `i > length - seg` (:209): only then does it print its epoch line.
    SEG = dict(seed=7, corpus_bytes=3001, epochs=2, B=4, S=12, N=64)
    out = run_reference(SEG["seed"], SEG["corpus_bytes"], SEG["epochs"], program="lstm_segment")
    read_line, epoch_lines, avg, gen, _ = parse(out)
    assert len(gen) >= SEG["epochs"] and all(len(g) == 1500 for g in gen[:SEG["epochs"]]), [len(g) for g in gen]
    libc.srand(1)
    positions = [[libc.rand() % (SEG["corpus_bytes"] - SEG["S"]) + SEG["S"] for _ in range(SEG["B"])] for _ in range(SEG["epochs"])]
    doc = dict(SEG, positions=positions, read_line=read_line, epoch_lines=epoch_lines[:SEG["epochs"]], avg_loss=avg[:SEG["epochs"]],
               generated_b64=[base64.b64encode(g).decode() for g in gen[:SEG["epochs"]]],
               how="oracle/_ref/lstm_segment_ref = unmodified OV/lstm_eigen_class_batch/lstm_segment.cc + lstm.h + oracle/eigen_shim, "
                   "REF_SEED=7, cwd holding a file named lstm.h = first 3001 bytes of R/alice29.txt")
    json.dump(doc, open(os.path.join(HERE, "ref_lstm_segment_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:SEG["epochs"]]))

    # the two remaining CPU snapshots: OV/lstm_eigen_opt (f32, "batched" with B = 1, S = 5, 2000 sampled characters) and
    # OV/lstm_eigen_class (double, N = 20, S = 50, B = 1, FULL numerical gradient check at the end of every epoch: slow, 1 epoch)
    OPT = dict(seed=321, corpus_bytes=2500, epochs=3, B=1, S=5, N=64)
    libc.srand(1)
    OPT["positions"] = [libc.rand() % (OPT["corpus_bytes"] - OPT["S"]) + OPT["S"]]
    out = run_reference(OPT["seed"], OPT["corpus_bytes"], OPT["epochs"], program="lstm_eigen_opt")
    read_line, epoch_lines, avg, gen, _ = parse(out)
    assert all(len(g) == 2000 for g in gen[:OPT["epochs"]])
    json.dump(dict(OPT, read_line=read_line, epoch_lines=epoch_lines[:OPT["epochs"]], avg_loss=avg[:OPT["epochs"]],
                   generated_b64=[base64.b64encode(g).decode() for g in gen[:OPT["epochs"]]],
                   how="oracle/_ref/lstm_eigen_opt_ref = unmodified OV/lstm_eigen_opt/lstm.cc + oracle/eigen_shim, REF_SEED=321"),
              open(os.path.join(HERE, "ref_lstm_eigen_opt_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:OPT["epochs"]]))
    CLS = dict(seed=55, corpus_bytes=2000, epochs=1, B=1, S=50, N=20)
    out = run_reference(CLS["seed"], CLS["corpus_bytes"], CLS["epochs"], program="lstm_eigen_class")
    read_line, epoch_lines, avg, gen, _ = parse(out)
    assert all(len(g) == 2500 for g in gen[:CLS["epochs"]])
    json.dump(dict(CLS, read_line=read_line, epoch_lines=epoch_lines[:CLS["epochs"]], avg_loss=avg[:CLS["epochs"]],
                   generated_b64=[base64.b64encode(g).decode() for g in gen[:CLS["epochs"]]],
                   how="oracle/_ref/lstm_eigen_class_ref = unmodified OV/lstm_eigen_class/lstm.cc + lstm.h + oracle/eigen_shim, REF_SEED=55"),
              open(os.path.join(HERE, "ref_lstm_eigen_class_run.json"), "w"), indent=1)
    print("\n".join(epoch_lines[:CLS["epochs"]]))


if __name__ == "__main__":
    sys.exit(main())
