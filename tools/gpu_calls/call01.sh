#!/bin/bash
# Round-2 GPU call 1: measure everything round 1 left "prepared, not yet run" + ncu captures of the kernels VERDICT names.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c01
mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
B="--no-cpu-baseline --no-e2e --steps 6 --warmup 3"

# (1) opt-in tests of the prepared kernels (split-K TN GEMM, window attention variant 2)
TA_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests -m gpu -x -q -k "split_k or window" > $O/pytest_unverified.log 2>&1
tail -3 $O/pytest_unverified.log

# (2) headline step on this box (reference point for the A/Bs) + in-situ trace
timeout 300 python bench.py $B --trace-kernels $O/trace_mlp.txt > $O/bench_mlp.json 2> $O/bench_mlp.err
# (3) PDL build A/B
TA_LIB_VARIANT=pdl TA_PDL=0 timeout 300 python bench.py $B > $O/bench_mlp_pdlbuild_off.json 2> $O/bench_mlp_pdlbuild_off.err
TA_LIB_VARIANT=pdl TA_PDL=1 timeout 300 python bench.py $B --trace-kernels $O/trace_mlp_pdl.txt > $O/bench_mlp_pdl.json 2> $O/bench_mlp_pdl.err
TA_LIB_VARIANT=pdl TA_PDL=1 timeout 900 python -m pytest tests/test_path_gpu.py -m gpu -x -q -k "not full_size" > $O/pytest_pdl.log 2>&1
tail -3 $O/pytest_pdl.log
# configs[2] literal: 8 clips per GPU
timeout 300 python bench.py $B --batch 8 > $O/bench_mlp_b8.json 2> $O/bench_mlp_b8.err
TA_LIB_VARIANT=pdl TA_PDL=1 timeout 300 python bench.py $B --batch 8 > $O/bench_mlp_b8_pdl.json 2> $O/bench_mlp_b8_pdl.err

# (4) QFormer: window attention variant 1 vs 2
TA_WINDOW_ATTN_VARIANT=1 timeout 300 python bench.py $B --projector qformer --trace-kernels $O/trace_qformer_v1.txt > $O/bench_qformer_v1.json 2> $O/bench_qformer_v1.err
TA_WINDOW_ATTN_VARIANT=2 timeout 300 python bench.py $B --projector qformer --trace-kernels $O/trace_qformer_v2.txt > $O/bench_qformer_v2.json 2> $O/bench_qformer_v2.err
# (5) LoRA: TN split-K off / on
TA_GEMM_TN_SPLITK=0 timeout 300 python bench.py $B --lora > $O/bench_lora_sk0.json 2> $O/bench_lora_sk0.err
TA_GEMM_TN_SPLITK=1 timeout 300 python bench.py $B --lora --trace-kernels $O/trace_lora_sk1.txt > $O/bench_lora_sk1.json 2> $O/bench_lora_sk1.err

# (6) ncu --set full: encoder attention (default variant), decoder GEMM epilogues
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd1 -s 1 -c 1 -o $O/ncu_attn_enc -f \
    python tools/prof_kernels.py attn_enc > $O/ncu_attn_enc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 4 -c 2 -o $O/ncu_lm_gemm -f \
    python tools/prof_kernels.py lm_gemm > $O/ncu_lm_gemm.log 2>&1
for f in $O/*.json; do echo "== $f"; head -c 600 $f; echo; done
ls -la $O
