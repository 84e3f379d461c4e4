#!/bin/bash
# Round 2, GPU call 16 (one B200): host-buffer path against chunk count + raw PCIe rates; MIXED kernel on the skewed
# stand-ins (sweep + ncu)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c16
mkdir -p $O
echo "== new tests"
timeout 600 python -m pytest tests/test_fused_halo_gpu.py tests/test_spmv_gpu.py -q -m gpu --timeout 300 -k "allgather or collective or ring_geometries" > $O/pytest_new.log 2>&1; tail -3 $O/pytest_new.log
timeout 600 python tools/e2e_chunks.py 8,16,32,4 c5 > $O/e2e_chunks_c5.jsonl 2> $O/e2e_err.log; cat $O/e2e_chunks_c5.jsonl; tail -2 $O/e2e_err.log
bash tools/gpu_call13.sh
