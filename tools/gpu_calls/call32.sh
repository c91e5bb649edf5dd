#!/bin/bash
# Round-2 GPU call 32: QFormer LayerNorm backward with shared-memory dw/db accumulators
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c32
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "add_layernorm or gelu_and_colsum" > $O/pytest_glue.log 2>&1; tail -3 $O/pytest_glue.log
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "qformer" > $O/pytest_qformer.log 2>&1; tail -3 $O/pytest_qformer.log
timeout 600 python bench.py --projector qformer --steps 6 --warmup 3 --no-other-configs --no-dp-parity --trace-kernels $O/trace_qformer.txt > $O/bench_qformer.json 2> $O/bench_qformer.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_qformer.json") if l.startswith("{")][-1]
print("qformer step", d["ms_per_step"], d["clocks"], d.get("loss"))
P
grep -n "add_layernorm\|window_attn\|gelu\|colsum" $O/trace_qformer.txt
