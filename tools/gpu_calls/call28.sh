#!/bin/bash
# Round-2 GPU call 28: D = rowsum(dO o O) fused into the o-projection dgrad GEMM epilogue (TA_EPI_BF16_ROWDOT)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c28
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "rowdot or attn_bwd or gemm" > $O/pytest_k.log 2>&1; tail -3 $O/pytest_k.log
timeout 1200 python -m pytest tests/test_path_gpu.py tests/test_ddp_gpu.py -m gpu -q > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
for v in 1 0 1 0; do
  TA_LM_FUSED_ATTN_DSUM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_dsum$v.txt > $O/bench_dsum$v.json 2> $O/bench_dsum$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_dsum$v.json") if l.startswith("{")][-1]
print("fused_dsum=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "attn_bwd_prep\|256, 8\|gemm2_kernel<256, 0" $O/trace_dsum$v.txt
done
