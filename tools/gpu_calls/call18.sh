#!/bin/bash
# Round-2 GPU call 18: ncu --set full of the pair GEMM vs cuBLAS (fc1 shape and 8192^3): tensor-pipe activity, L2 throughput, launch shapes
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c18
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -o $O/ncu_gemm_cmp python tools/prof_kernels.py gemm_cmp > $O/ncu_gemm_cmp.log 2>&1; tail -3 $O/ncu_gemm_cmp.log
ncu -i $O/ncu_gemm_cmp.ncu-rep --page raw --csv > $O/gemm_cmp_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/gemm_cmp_raw.csv | cut -c1-200
