"""Golden OUTPUT of the reference's own source: run oracle/_ref/lstm_ref and record what it prints.

    make -C oracle ref && python tests/golden/make_ref_run.py        (in the build container)

oracle/_ref/lstm_ref is the UNMODIFIED /root/reference/lstm.cc compiled against oracle/eigen_shim (the real Eigen
is not installed; see oracle/Makefile) with std::random_device replaced by a seed counter (REF_SEED).  The program
has no arguments: it reads "alice29.txt" from its working directory and runs 1000 epochs, so it is started in a
scratch directory holding the first CORPUS_BYTES bytes of R/alice29.txt (= tests/golden/alice29_head.bin[:3000])
and stopped after EPOCHS epoch reports.

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
SEED, CORPUS_BYTES, EPOCHS = 1234, 3000, 4
END = b"| Generated text END ************"


def run_reference(seed=SEED, corpus_bytes=CORPUS_BYTES, epochs=EPOCHS, timeout=120):
    """Start lstm_ref on the truncated corpus, read its stdout until `epochs` samples were printed, stop it."""
    text = open(os.path.join(HERE, "alice29_head.bin"), "rb").read()[:corpus_bytes]
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "alice29.txt"), "wb").write(text)
        p = subprocess.Popen([BIN], cwd=d, stdout=subprocess.PIPE, env=dict(os.environ, REF_SEED=str(seed)))
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


if __name__ == "__main__":
    sys.exit(main())
