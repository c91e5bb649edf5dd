#!/bin/bash
# Round-2 GPU call 44: pairwise bf16 rounding in the GEMM epilogues (F2F.BF16.F32 off the XU pipe)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c44
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or swiglu or rowdot" > $O/pytest_gemm.log 2>&1; tail -3 $O/pytest_gemm.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; head -8 $O/time_ffn.log
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_mlp.txt > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("step", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"), round(d["roofline"]["achieved"]))
P
sed -n 3,16p $O/trace_mlp.txt
