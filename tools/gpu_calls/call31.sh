#!/bin/bash
# Round-2 GPU call 31: norm kernels over warps-per-block; step with the best setting
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c31
mkdir -p $O
timeout 300 python tools/sweep_norms.py > $O/sweep_norms.log 2>&1; cat $O/sweep_norms.log
