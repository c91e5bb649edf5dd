#!/bin/bash
# Round-2 GPU call 17: L2 prefetch in the SwiGLU-backward stash agent; per-kernel energy profile (power-capped step)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c17
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "swiglu" > $O/pytest_gemm.log 2>&1; tail -3 $O/pytest_gemm.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; cat $O/time_ffn.log
timeout 600 python tools/energy_profile.py --seconds 1.5 > $O/energy_profile.txt 2> $O/energy_profile.err; cat $O/energy_profile.txt; tail -3 $O/energy_profile.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("step", d["ms_per_step"], d["clocks"], d.get("loss"))
P
