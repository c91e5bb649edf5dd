#!/bin/bash
# Round-2 GPU call 14: default attention mode 14, no mma.sync fallback, device prompt assembly; LM attention timing; bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c14
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1
tail -8 $O/pytest_gpu.log
timeout 300 python tools/time_lm_attn.py > $O/time_lm_attn.log 2>&1; cat $O/time_lm_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd_kernel -s 1 -c 1 -o $O/ncu_attn_bwd -f python tools/prof_kernels.py attn_lm > $O/ncu_attn_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attn_tc_fwd_kernel" -s 1 -c 1 -o $O/ncu_attn_lmfwd -f python tools/prof_kernels.py attn_lm > $O/ncu_attn_lmfwd.log 2>&1
(time timeout 900 python bench.py --steps 10 --warmup 3) > $O/bench.json 2> $O/bench.err
head -c 400 $O/bench.json; echo; tail -3 $O/bench.err
