#!/bin/bash
# Round-2 GPU call 41: which kernel makes the decoder backward depend on the batch size?
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c41
mkdir -p $O
timeout 900 python tools/batch_invariance.py 16 30 > $O/inv_16x30.log 2>&1; tail -9 $O/inv_16x30.log
