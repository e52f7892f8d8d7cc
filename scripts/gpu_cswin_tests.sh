#!/bin/bash
# GA-CSWin GPU parity tests (stripe attention, CSWinBlock, whole model), each group in its own process.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ga_cswin.py -q -m gpu -k "stripe_attention" -p no:cacheprovider > gpurun_out/t_cs_attn.log 2>&1; echo "attn rc=$?"
timeout 900 python -m pytest tests/test_ga_cswin.py -q -m gpu -k "test_block" -p no:cacheprovider > gpurun_out/t_cs_blk.log 2>&1; echo "blk rc=$?"
timeout 1200 python -m pytest tests/test_ga_cswin.py -q -m gpu -k "test_model" -p no:cacheprovider > gpurun_out/t_cs_model.log 2>&1; echo "model rc=$?"
for f in t_cs_attn t_cs_blk t_cs_model; do echo "== $f"; tail -${TAILN:-12} gpurun_out/$f.log; done
