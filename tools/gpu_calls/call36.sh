#!/bin/bash
# Round-2 GPU call 36: batch invariance of the forward (why does the same clip's loss depend on the shard it sits in?)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c36
mkdir -p $O
timeout 600 python tools/batch_invariance.py 8 30 > $O/inv_8x30.log 2>&1; tail -12 $O/inv_8x30.log
timeout 600 python tools/batch_invariance.py 64 30 > $O/inv_64x30.log 2>&1; tail -12 $O/inv_64x30.log
