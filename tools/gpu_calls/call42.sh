#!/bin/bash
# Round-2 GPU call 42: final evidence -- ncu --set full of the decoder-attention kernels and the ROWDOT GEMM on the final library; full suite; bench
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c42
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_fwd6|attn_tc_bwd2" -o $O/ncu_lm_attn python tools/prof_kernels.py attn_lm > $O/ncu_lm_attn.log 2>&1; tail -1 $O/ncu_lm_attn.log
ncu -i $O/ncu_lm_attn.ncu-rep --page raw --csv > $O/lm_attn_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/lm_attn_raw.csv > $O/lm_attn_summary.txt; grep -A6 "==" $O/lm_attn_summary.txt | cut -c1-130
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
python - <<P
import json
d=[json.loads(l) for l in open("$O/bench.json") if l.startswith("{")][-1]
print("bench", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"], d.get("loss"))
print({k:(v.get("ms_per_step")) for k,v in (d.get("other_configs") or {}).items()})
print({k:(round(v.get("tflops")), round(v.get("frac"),3)) for k,v in d["roofline"]["qwen3_ffn"].items() if isinstance(v, dict)}, round(d["roofline"]["achieved"]), round(d["roofline"]["frac"],3))
print(d["parity"]["ce_loss_delta"], d["parity"]["batch_sample"]["ce_loss_delta"], d["cpu_baseline"])
P
