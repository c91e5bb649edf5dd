#!/bin/bash
# Round-2 GPU call 20: cooperative 64-column in-place SwiGLU-backward epilogue + relaxed accumulator-consumed arrive (all pair GEMMs);
# QFormer glue kernels after the modulo hoist / two-stage dw,db reduction
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c20
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "gemm or swiglu or add_layernorm or gelu_and_colsum" > $O/pytest_kernels.log 2>&1; tail -3 $O/pytest_kernels.log
timeout 300 python tools/time_ffn.py > $O/time_ffn.log 2>&1; cat $O/time_ffn.log
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --trace-kernels $O/trace_mlp.txt > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --projector qformer --steps 6 --warmup 3 --no-other-configs --no-dp-parity --trace-kernels $O/trace_qformer.txt > $O/bench_qformer.json 2> $O/bench_qformer.err
python - <<P
import json
for f in ("bench","bench_qformer"):
    d=[json.loads(l) for l in open("$O/%s.json"%f) if l.startswith("{")][-1]
    print(f, d["ms_per_step"], d["clocks"], d.get("loss"))
P
grep -n "add_layernorm\|gelu\|colsum\|256, 6\|256, 3\|256, 5" $O/trace_qformer.txt $O/trace_mlp.txt
