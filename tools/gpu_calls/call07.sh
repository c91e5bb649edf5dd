#!/bin/bash
# Round-2 GPU call 7: persistent attention with one MMA-issuing thread per group
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c07
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "test_attn_fwd" > $O/pytest_attn.log 2>&1
tail -4 $O/pytest_attn.log
timeout 300 python tools/time_attn.py 2 6 7 8 12 13 > $O/time_attn.log 2>&1
cat $O/time_attn.log
timeout 300 python tools/attn_trace.py 7 8 > $O/attn_trace.log 2>&1
cat $O/attn_trace.log
