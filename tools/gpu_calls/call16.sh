#!/bin/bash
# Round-2 GPU call 16: in-place TMA epilogue of the SwiGLU-backward GEMM -- parity, per-GEMM timing, step A/B (and the tail-split switch again)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c16
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "swiglu or gemm" > $O/pytest_gemm.log 2>&1; tail -5 $O/pytest_gemm.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; cat $O/time_ffn.log
for v in "1 0" "0 0" "1 1"; do set -- $v
  TA_GEMM_SWIGLU_BWD_TMA=$1 TA_GEMM_TAIL_SPLIT=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity > $O/bench_tma$1_tail$2.json 2> $O/bench_tma$1_tail$2.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_tma$1_tail$2.json") if l.startswith("{")][-1]
print("swiglu_bwd_tma=$1 tail_split=$2", d["ms_per_step"], d["clocks"], d.get("loss"))
P
done
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
