#!/bin/bash
# Round 2, GPU call 4 (one B200): persistent ring form of the staged-x kernels; plan choices on SuiteSparse shapes.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c4
mkdir -p $O
echo "== parity"
timeout 1200 python -m pytest tests/test_spmv_gpu.py tests/test_analysis_gpu.py tests/test_fused_halo_gpu.py -q -m gpu --timeout 600 -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -25 $O/pytest_quick.log
echo "== ring sweeps"
timeout 600 python tools/sweep.py --workloads c5s --tiles 0,1792,2304,2816 --vecdivs 0,8 --xflags 0 --ring 2x0,1x0,2x2,1x3,3x0 --reps 30 > $O/sweep_c5s_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c5s --tiles 0 --xflags 33554432,262144 --reps 30 >> $O/sweep_c5s_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0,1024,2048,3072 --xflags 0 --ring 2x0,1x0,3x0,4x0,2x3 --reps 50 > $O/sweep_c2_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0 --xflags 33554432,262144 --reps 50 >> $O/sweep_c2_ring.jsonl 2>&1
cat $O/sweep_c5s_ring.jsonl $O/sweep_c2_ring.jsonl | cut -c1-210
echo "== plan choices on the shapes where the default lost to cuSPARSE"
for w in ss:Ga41As41H72 ss:vas_stokes_2M ss:TSOPF_RS_b2383 ss:dielFilterV3real ss:Hardesty3; do
  timeout 300 python tools/sweep.py --workloads $w --tiles 0,1024,2048,4096 --xflags 0,64,128 --cusparse --reps 30 >> $O/sweep_ss.jsonl 2>&1
done
cat $O/sweep_ss.jsonl | cut -c1-210
echo "== ncu ring kernel on c5s"
timeout 300 python tools/profile_one.py c5s > $O/plain_c5s.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c5s_ring python tools/profile_one.py c5s > $O/ncu_c5s.log 2>&1
cat $O/plain_c5s.log
ls -la $O
