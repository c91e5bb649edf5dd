#!/bin/bash
# Round-2 GPU call 37-38 (2 GPUs): dp_loss_delta diagnosis (parameter drift across replicas, rank-0 shard re-run)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c38
mkdir -p $O
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 4 --warmup 3 --no-other-configs --no-e2e) > $O/bench_n2.json 2> $O/bench_n2.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_n2.json") if l.startswith("{")][-1]
print("N=2", d["ms_per_step"], "dp", d["parity"]["dp"])
P
tail -3 $O/bench_n2.err
