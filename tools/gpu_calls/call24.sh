#!/bin/bash
# Round-2 GPU call 24: decoder attention -- forward on 64-key tiles with two CTAs per SM (fwd6), backward with dP issued under the dQ
# read-out and LSE / D prefetched one iteration ahead
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c24
mkdir -p $O
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "attn_fwd_decoder or attn_bwd or (attn_fwd and 128)" > $O/pytest_attn.log 2>&1; tail -4 $O/pytest_attn.log
timeout 300 python tools/time_lm_attn.py > $O/time_lm_attn.log 2>&1; cat $O/time_lm_attn.log
timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -q -x > $O/pytest_path.log 2>&1; tail -3 $O/pytest_path.log
for v in 1 0; do
  TA_ATTN_TC_LM=$v timeout 600 python bench.py --steps 8 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline --trace-kernels $O/trace_lm$v.txt > $O/bench_lm$v.json 2> $O/bench_lm$v.err
  python - <<P
import json
d=[json.loads(l) for l in open("$O/bench_lm$v.json") if l.startswith("{")][-1]
print("attn_tc_lm=$v", d["ms_per_step"], d["clocks"]["sm_mhz"], d.get("loss"))
P
  grep -n "attn_tc_fwd6\|attn_tc_fwd_kernel<128\|attn_tc_bwd" $O/trace_lm$v.txt
done
