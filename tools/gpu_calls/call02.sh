#!/bin/bash
# Round-2 GPU call 2: full GPU test suite with the new parity tests (dropout, full-size greedy ids, surface, DDP), the new bench line.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c02
mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
(time timeout 1500 python -m pytest tests -m gpu -q -s --durations=15) > $O/pytest_gpu.log 2>&1
tail -40 $O/pytest_gpu.log
(time timeout 900 python bench.py --steps 10 --warmup 3) > $O/bench.json 2> $O/bench.err
tail -c 3000 $O/bench.json; tail -5 $O/bench.err
(time timeout 600 python bench.py --impl reference --steps 2 --warmup 1) > $O/bench_ref.json 2> $O/bench_ref.err
tail -c 1500 $O/bench_ref.json
ls -la $O
