#!/bin/bash
# Round 2, GPU call 10 (one B200): full GPU test suite, full bench line + reference arm, ncu launch list of the bench,
# ncu --set full of the ring kernels on C2 and on the full-size C5 matrix, L2-policy probe on small matrices
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c10
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,memory.total --format=csv > $O/smi.txt 2>&1
echo "== full GPU tests"
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log; tail -5 $O/pytest_full.log
echo "== bench N=1"
timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; rc=$?; echo "rc=$rc"
tail -3 $O/bench_n1.err
echo "== reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"
cat $O/bench_ref.json | cut -c1-600
echo "== small matrices: L2 policy of the streams"
timeout 300 python tools/sweep.py --workloads ss:Ga41As41H72,ss:largebasis,ss:vas_stokes_2M,ss:TSOPF_RS_b2383,ss:boneS10,ss:Hardesty3 --tiles 0 --xflags 0,134217728 --reps 50 --cusparse > $O/sweep_small_l2.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2c10/sweep_small_l2.jsonl"):
    try:
        d = json.loads(l)
        print(d["workload"], d["tile"], d["flags"], d.get("ring_used"), d["ms"], d["gbs"], d.get("xstage"), d.get("ms_cusparse")) if "ms" in d else print(d)
    except Exception:
        print("??", l[:160])
PY
if [ $rc -eq 0 ]; then
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv \
  python bench.py --steps 20 --warmup 3 --no-other-configs --no-context --no-cpu > $O/bench_under_ncu.log 2>&1; echo "rc=$?"
fi
echo "== ncu --set full: C2 and C5 (full size)"
timeout 300 python tools/profile_one.py c2 > $O/plain_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c2_ring python tools/profile_one.py c2 > $O/ncu_c2.log 2>&1
cat $O/plain_c2.log
timeout 300 python tools/profile_one.py c5 > $O/plain_c5.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c5_ring python tools/profile_one.py c5 > $O/ncu_c5.log 2>&1
cat $O/plain_c5.log; tail -3 $O/ncu_c5.log
ls -la $O
