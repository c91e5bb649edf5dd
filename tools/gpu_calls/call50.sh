#!/bin/bash
# Round-2 GPU call 50: repeat the in-step A/B of the residual epilogue (0,1,0,1)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c50
mkdir -p $O
for v in 0 1 0 1; do
  TA_GEMM_RESID_TMA=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --no-e2e > $O/bench_resid$v.json 2> $O/bench_resid$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_resid$v.json") if l.startswith("{")][-1]
print("resid_tma=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
done
