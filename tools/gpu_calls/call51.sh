#!/bin/bash
# Round-2 GPU call 51: full GPU suite on the final binary
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c51
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
