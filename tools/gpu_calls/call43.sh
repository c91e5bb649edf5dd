#!/bin/bash
# Round-2 GPU call 43: LoRA / unfrozen-decoder recipes -- h kept per layer, LoRA gradient buffers cleared once per step
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c43
mkdir -p $O
timeout 1200 python -m pytest tests/test_path_gpu.py -m gpu -q -k "lora or unfrozen or decoder_grad or surface" > $O/pytest_lora.log 2>&1; tail -3 $O/pytest_lora.log
timeout 600 python bench.py --lora --steps 6 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_lora.txt > $O/bench_lora.json 2> $O/bench_lora.err
timeout 600 python bench.py --train-lm --steps 5 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline > $O/bench_trainlm.json 2> $O/bench_trainlm.err
python - <<P
import json
for f in ("bench_lora","bench_trainlm"):
    d=[json.loads(l) for l in open("$O/%s.json"%f) if l.startswith("{")][-1]
    print(f, d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"], d.get("loss"))
P
grep -n "swiglu_h\|Memset\|Memcpy DtoD\|gemm2_kernel<128, 0\|true, true" $O/trace_lora.txt
