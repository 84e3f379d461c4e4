#!/bin/bash
# Round 2, GPU call 5 (one B200): ring kernels for real (call 4 silently fell back: occupancy query before the smem limit was raised)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c5
mkdir -p $O
echo "== ring smoke (tight time limit)"
timeout 240 python -m pytest tests/test_spmv_gpu.py -q -m gpu --timeout 90 -x -k "ring_geometries or windows" > $O/pytest_ring.log 2>&1; rc=$?; echo "rc=$rc" >> $O/pytest_ring.log; tail -15 $O/pytest_ring.log
if [ $rc -ne 0 ]; then echo "ring smoke failed: stopping"; exit 0; fi
echo "== parity"
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_analysis_gpu.py tests/test_fused_halo_gpu.py -q -m gpu --timeout 300 -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -15 $O/pytest_quick.log
echo "== ring sweeps"
timeout 600 python tools/sweep.py --workloads c5s --tiles 0,1792,2304,2816 --vecdivs 0,8 --xflags 0 --ring 2x0,1x0,2x2,1x3,3x0 --reps 30 > $O/sweep_c5s_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c5s --tiles 0 --xflags 33554432,262144 --reps 30 >> $O/sweep_c5s_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0,1024,2048 --xflags 0 --ring 2x0,1x0,3x0,4x0,2x3 --reps 50 > $O/sweep_c2_ring.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0 --xflags 33554432,262144 --reps 50 >> $O/sweep_c2_ring.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/r2c5/sweep_c5s_ring.jsonl", "gpurun_out/r2c5/sweep_c2_ring.jsonl"):
    for l in open(f):
        try:
            d = json.loads(l)
            print(d["workload"], d["tile"], d["vec_div"], d["flags"], d.get("ring"), d.get("ring_used"), d["ms"], d["gbs"], d.get("xstage"), d["smem"]) if "ms" in d else print(d)
        except Exception:
            print("??", l[:160])
PY
echo "== ncu ring kernel on c5s"
timeout 300 python tools/profile_one.py c5s > $O/plain_c5s.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_c5s_ring python tools/profile_one.py c5s > $O/ncu_c5s.log 2>&1
cat $O/plain_c5s.log
echo "== bench (headline + e2e only)"
timeout 600 python bench.py --quick --no-cpu > $O/bench_quick.json 2> $O/bench_quick.err; echo "rc=$?"; tail -c 600 $O/bench_quick.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c5/bench_quick.json"))
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["verified"], d["config"]["tile_nnz"], d["e2e"].get("value"))
for k, v in d.get("iterated", {}).items():
    print(" iter", k, v.get("ms_per_iter"), v.get("x_checksum_first_16th"), v.get("launches_per_iteration"), v.get("note", ""))
PY
ls -la $O
