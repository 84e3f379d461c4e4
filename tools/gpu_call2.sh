#!/bin/bash
# Round 2, GPU call 2 (one B200): staged-x form (parity, sweeps, ncu), evict_last gathers, reference harness, bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c2
mkdir -p $O
echo "== parity (analysis arrays, every option set, fused loop, reference CLI + harness)"
timeout 1200 python -m pytest tests/test_analysis_gpu.py tests/test_spmv_gpu.py tests/test_fused_halo_gpu.py tests/test_reference_cli_gpu.py \
  -q -m gpu --timeout 600 -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -25 $O/pytest_quick.log
echo "== sweeps"
timeout 600 python tools/sweep.py --workloads c5s --tiles 0,1792,2304,2816,3328 --vecdivs 0,8 --xflags 0,262144,16777216 --reps 30 > $O/sweep_c5s.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0,1024,1536,2048 --xflags 0,262144 --reps 50 > $O/sweep_c2.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c3 --tiles 0 --xflags 0,134217728 --reps 20 > $O/sweep_c3.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c4 --tiles 0 --xflags 0,134217728,8388608 --reps 20 > $O/sweep_c4.jsonl 2>&1
cat $O/sweep_c5s.jsonl $O/sweep_c2.jsonl $O/sweep_c3.jsonl $O/sweep_c4.jsonl | cut -c1-230
echo "== bench default"
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "rc=$?"; tail -c 1000 $O/bench_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c2/bench_n1.json"))
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["verified"], d["e2e"].get("value"))
for k, v in d.get("iterated", {}).items():
    print(" iter", k, v.get("ms_per_iter"), v.get("x_checksum_first_16th"), v.get("note", ""))
for k, v in d.get("other_configs", {}).items():
    print(" other", k, v.get("ms"), v.get("verified"), v.get("error", ""))
PY
echo "== ncu (staged-x kernel on c5s, full suite of sections, source)"
timeout 300 python tools/profile_one.py c5s > $O/plain_c5s.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c5s_xs python tools/profile_one.py c5s > $O/ncu_c5s.log 2>&1
timeout 300 python tools/profile_one.py c2 > $O/plain_c2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c2_xs python tools/profile_one.py c2 > $O/ncu_c2.log 2>&1
cat $O/plain_c5s.log $O/plain_c2.log
echo "== full-size parity"
timeout 900 python -m pytest tests/test_spmv_gpu.py -q -m gpu --timeout 600 -k "full_size" > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log; tail -8 $O/pytest_full.log
ls -la $O
