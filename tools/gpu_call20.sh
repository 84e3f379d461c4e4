#!/bin/bash
# Round 2, GPU call 20 (one B200): compute-sanitizer (memcheck, then racecheck) over the ring / staged-x / halo tests
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c20
mkdir -p $O
SEL='ring_geometries or collective or rowptr_window or (golden and not full_size)'
timeout 300 python -m pytest tests/test_spmv_gpu.py tests/test_fused_halo_gpu.py -q -m gpu --timeout 200 -x -k "$SEL or allgather or single_rank" > $O/plain.log 2>&1; rc=$?; tail -2 $O/plain.log
if [ $rc -ne 0 ]; then echo "plain run failed"; exit 0; fi
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/memcheck.log \
  python -m pytest tests/test_spmv_gpu.py tests/test_fused_halo_gpu.py -q -m gpu --timeout 800 -x -k "$SEL or allgather or single_rank" > $O/memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 $O/memcheck_pytest.log; tail -4 $O/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/racecheck.log \
  python -m pytest tests/test_spmv_gpu.py -q -m gpu --timeout 800 -x -k "ring_geometries" > $O/racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -3 $O/racecheck_pytest.log; tail -6 $O/racecheck.log
