#!/bin/bash
# Round-2 GPU call 39 (8 GPUs): scaling + same-global-batch parity at N = 8 and N = 4 on the final tree
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c39
mkdir -p $O
for n in 8 4; do
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 6 --warmup 3 --no-other-configs) > $O/bench_n$n.json 2> $O/bench_n$n.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_n$n.json") if l.startswith("{")][-1]
print("N=$n value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"], "dp", {k: d["parity"]["dp"].get(k) for k in ("dp_loss_delta", "grad_rel_err_vs_single_gpu", "param_max_abs_diff_across_ranks", "loss_dp", "loss_single_gpu")})
P
tail -2 $O/bench_n$n.err
done
