#!/bin/bash
# Round-2 GPU call 35 (2 GPUs): deterministic double-precision loss reduction -- CE / path tests, same-global-batch parity at N = 2
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c35
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_path_gpu.py -m gpu -q -k "ce or cross or loss or dropout or small or full_size or smoke or batch_properties" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 4 --warmup 3 --no-other-configs) > $O/bench_n2.json 2> $O/bench_n2.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_n2.json") if l.startswith("{")][-1]
print("N=2 value", d["value"], d["ms_per_step"], "dp", d["parity"]["dp"])
P
tail -3 $O/bench_n2.err
