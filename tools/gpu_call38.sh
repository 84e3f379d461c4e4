#!/bin/bash
# Round 2, GPU call 38 (one B200): ncu --set full of the kernel instantiation the headline times (k_spmv_ring<MEDIUM, HALO>,
# beta = 0, launched from the CUDA graph of the power loop)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c38
mkdir -p $O
CMD="python bench.py --quick --steps 20 --warmup 3 --no-e2e --no-cpu --no-other-configs"
timeout 600 $CMD > $O/plain.json 2> $O/plain.err; rc=$?; echo "plain rc=$rc"
if [ $rc -eq 0 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv_ring -s 12 -c 1 -f -o $O/ncu_c5_halo_ring $CMD > $O/ncu.log 2>&1; echo "ncu rc=$?"
tail -3 $O/ncu.log | cut -c1-200
fi
ls -la $O
