#!/bin/bash
# Round 2, GPU call 3 (one B200): staged-x kernels with predicated-PTX slots, TMA gather4 probe, selector study.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c3
mkdir -p $O
echo "== parity"
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_analysis_gpu.py tests/test_fused_halo_gpu.py -q -m gpu --timeout 600 -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -15 $O/pytest_quick.log
echo "== sweeps"
timeout 600 python tools/sweep.py --workloads c5s --tiles 0,1792,2304,2816,3328 --vecdivs 0,8 --xflags 0,262144,16777216 --reps 30 > $O/sweep_c5s.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0,1024,1536,2048 --xflags 0,262144 --reps 50 > $O/sweep_c2.jsonl 2>&1
cat $O/sweep_c5s.jsonl $O/sweep_c2.jsonl | cut -c1-200
echo "== ncu (staged-x kernel on c5s)"
timeout 300 python tools/profile_one.py c5s > $O/plain_c5s.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c5s_xs python tools/profile_one.py c5s > $O/ncu_c5s.log 2>&1
cat $O/plain_c5s.log
echo "== TMA gather4 probe"
timeout 600 python tools/gather_bound.py --gather4 --tables 2500000,5000000,10000000 --reps 5 > $O/gather4.jsonl 2> $O/gather4.err; echo "rc=$?"; cat $O/gather4.jsonl | cut -c1-300; tail -5 $O/gather4.err
echo "== selector study"
timeout 900 python -m pytest tests/test_selector_study_gpu.py -q -m gpu --timeout 800 > $O/pytest_selector.log 2>&1; echo "rc=$?" >> $O/pytest_selector.log; tail -15 $O/pytest_selector.log
cp gpurun_out/selector_study*.json $O/ 2>/dev/null
ls -la $O
