#!/bin/bash
# Round-2 GPU call 33: DevicePrefetcher (H2D of batch i+1 under step i) -- test + bench e2e
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c33
mkdir -p $O
timeout 600 python -m pytest tests/test_path_gpu.py -m gpu -q -x -k "prefetcher or surface" > $O/pytest_prefetch.log 2>&1; tail -3 $O/pytest_prefetch.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("value", d["value"], d["ms_per_step"], "e2e", d["e2e"], d["clocks"])
P
tail -3 $O/bench.err
