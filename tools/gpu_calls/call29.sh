#!/bin/bash
# Round-2 GPU call 29: dQ accumulator zero-filled from the ROWDOT GEMM epilogue (no memset)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c29
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "rowdot or attn_bwd or gemm" > $O/pytest_k.log 2>&1; tail -3 $O/pytest_k.log
timeout 1200 python -m pytest tests/test_path_gpu.py tests/test_ddp_gpu.py -m gpu -q > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
for v in 1 0 1 0; do
  TA_LM_FUSED_ATTN_DSUM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_dsum$v.txt > $O/bench_dsum$v.json 2> $O/bench_dsum$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_dsum$v.json") if l.startswith("{")][-1]
print("fused_dsum+zero=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "attn_bwd_prep\|256, 8\|Memset" $O/trace_dsum$v.txt
done
