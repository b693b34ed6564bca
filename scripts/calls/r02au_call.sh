#!/bin/bash
# 1 GPU: the tree as committed last: smoke + the whole GPU suite
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee $OUT/r02au_smoke.txt
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee $OUT/r02au_pytest_gpu.txt
