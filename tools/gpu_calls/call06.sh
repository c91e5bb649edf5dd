#!/bin/bash
# Round-2 GPU call 6: timeline of the persistent attention kernel + ragged generate tests + ncu pipes for the integer-packed mode 2
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c06
mkdir -p $O
timeout 300 python tools/attn_trace.py 7 8 6 > $O/attn_trace.log 2>&1
cat $O/attn_trace.log
timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -k "ragged or greedy or kv_cache or generate" > $O/pytest_gen.log 2>&1
tail -15 $O/pytest_gen.log
TA_ATTN_TC=2 timeout 600 ncu --set full --clock-control none -k regex:attn_tc_fwd1 -s 1 -c 1 -o $O/ncu_attn2 -f python tools/prof_kernels.py attn_enc > $O/ncu_attn2.log 2>&1
ls -la $O
