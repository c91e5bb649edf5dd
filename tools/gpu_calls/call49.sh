#!/bin/bash
# Round-2 GPU call 49: in-place TMA epilogue for the bf16-residual GEMMs (encoder o-projection / fc2) -- parity, timing, step A/B
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c49
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm" > $O/pytest_gemm.log 2>&1; tail -3 $O/pytest_gemm.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; tail -4 $O/time_ffn.log
for v in 1 0; do
  TA_GEMM_RESID_TMA=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --no-e2e > $O/bench_resid$v.json 2> $O/bench_resid$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_resid$v.json") if l.startswith("{")][-1]
print("resid_tma=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
done
