#!/bin/bash
# Round-2 GPU call 23: evidence on the final library -- ncu --set full of the top kernels, ncu launch list of one bench step, smoke()
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c23
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm2_kernel|attn_tc" -o $O/ncu_top python tools/prof_kernels.py gemm down swiglu_bwd attn_enc attn_lm > $O/ncu_top.log 2>&1; tail -2 $O/ncu_top.log
ncu -i $O/ncu_top.ncu-rep --page raw --csv > $O/top_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/top_raw.csv > $O/top_summary.txt; grep -c "==" $O/top_summary.txt
TA_PROFILE_STEP=1 timeout 1200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-other-configs --no-dp-parity --no-cpu-baseline > $O/bench_under_ncu.log 2>&1; tail -1 $O/bench_under_ncu.log | cut -c1-200
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1; head -30 $O/launches_summary.txt
