#!/bin/bash
# Round-2 GPU call 40: batch (in)variance of the BACKWARD -- whole batch vs sum over shards, per projector tensor; new determinism test
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c40
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "deterministic" > $O/pytest_det.log 2>&1; tail -2 $O/pytest_det.log
timeout 900 python tools/batch_invariance.py 32 30 > $O/inv_32x30.log 2>&1; tail -12 $O/inv_32x30.log
