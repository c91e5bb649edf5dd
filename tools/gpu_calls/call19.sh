#!/bin/bash
# Round-2 GPU call 19: QFormer glue kernels (fused dropout+residual+LayerNorm, GELU, column sums) -- unit + path tests, qformer step;
# ncu of the SwiGLU-backward GEMM with the in-place TMA epilogue
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c19
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "add_layernorm or gelu_and_colsum" > $O/pytest_glue.log 2>&1; tail -5 $O/pytest_glue.log
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "qformer or mixture or surface" > $O/pytest_qformer.log 2>&1; tail -8 $O/pytest_qformer.log
timeout 600 python bench.py --projector qformer --steps 6 --warmup 3 --no-other-configs --no-dp-parity --trace-kernels $O/trace_qformer.txt > $O/bench_qformer.json 2> $O/bench_qformer.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_qformer.json") if l.startswith("{")][-1]
print("qformer step", d["ms_per_step"], d["clocks"], d.get("loss"))
P
head -45 $O/trace_qformer.txt
timeout 600 ncu --set full --clock-control none --import-source on -o $O/ncu_swiglu_bwd python tools/prof_kernels.py swiglu_bwd > $O/ncu_swiglu_bwd.log 2>&1; tail -2 $O/ncu_swiglu_bwd.log
ncu -i $O/ncu_swiglu_bwd.ncu-rep --page raw --csv > $O/swb_raw.csv 2>/dev/null
ncu -i $O/ncu_swiglu_bwd.ncu-rep --page source --csv --launch-skip 2 --launch-count 1 > $O/swb_src.csv 2>/dev/null
python tools/ncu_summary.py $O/swb_raw.csv | grep -A14 "gemm2" | cut -c1-160
