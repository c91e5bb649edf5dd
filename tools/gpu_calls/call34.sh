#!/bin/bash
# Round-2 GPU call 34 (2 GPUs): final check of the driver's commands on the final tree -- full GPU suite, smoke(), bench at N = 2 and the reference arm
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c34
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 6 --warmup 3) > $O/bench_n2.json 2> $O/bench_n2.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_n2.json") if l.startswith("{")][-1]
print("N=2 value", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"], "dp", d["parity"]["dp_loss_delta"], d["parity"]["dp"]["grad_rel_err_vs_single_gpu"])
print({k:(v.get("ms_per_step")) for k,v in (d.get("other_configs") or {}).items()})
P
tail -4 $O/bench_n2.err
