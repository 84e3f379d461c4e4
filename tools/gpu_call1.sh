#!/bin/bash
# Round 2, GPU call 1 (one B200): parity suite, C4 / C3 sweeps of the new direct kernel, ncu captures, bench smoke + run.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c1
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi.txt 2>&1
echo "== quick parity first (direct kernel, fused loop, cache)"
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_fused_halo_gpu.py -q -m gpu -x --timeout 600 \
  -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -5 $O/pytest_quick.log
echo "== sweeps"
timeout 600 python tools/sweep.py --workloads c4 --tiles 1024,2048,4096 --xflags 0,8388608 --cusparse --reps 30 > $O/sweep_c4.jsonl 2>&1
timeout 600 python tools/sweep.py --workloads c3 --tiles 0,1024,2048 --xflags 0,4,64,8388672,68 --cusparse --reps 20 > $O/sweep_c3.jsonl 2>&1
tail -20 $O/sweep_c4.jsonl $O/sweep_c3.jsonl
echo "== bench smoke (small grid)"
timeout 600 python bench.py --grid 128 --steps 10 --warmup 3 --no-other-configs > $O/bench_smoke.json 2> $O/bench_smoke.err; echo "rc=$?"; tail -c 1500 $O/bench_smoke.err
echo "== full parity suite"
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log; tail -15 $O/pytest_full.log
echo "== bench default"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; tail -c 1000 $O/bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
echo "== ncu"
for w in c4 c3 c5s; do
  timeout 300 python tools/profile_one.py $w > $O/plain_$w.log 2>&1 && \
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_$w python tools/profile_one.py $w > $O/ncu_$w.log 2>&1
  tail -2 $O/plain_$w.log
done
ls -la $O
