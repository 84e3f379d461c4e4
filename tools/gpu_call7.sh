#!/bin/bash
# Round 2, GPU call 7 (one B200): staged-x ring with a dedicated producer warp and full / empty barriers per stage
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c7
mkdir -p $O
echo "== ring smoke"
timeout 240 python -m pytest tests/test_spmv_gpu.py -q -m gpu --timeout 90 -x -k "ring_geometries or windows or golden" > $O/pytest_ring.log 2>&1; rc=$?; echo "rc=$rc" >> $O/pytest_ring.log; tail -6 $O/pytest_ring.log
if [ $rc -ne 0 ]; then echo "ring smoke failed: stopping"; exit 0; fi
echo "== sweeps"
timeout 600 python tools/sweep.py --workloads c5s --tiles 0,1280,1664,2304,2816 --xflags 0 --ring 2x0,3x0,1x0 --reps 30 > $O/sweep_c5s.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c5s --tiles 0,2816 --xflags 33554432,262144 --reps 30 >> $O/sweep_c5s.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0,1024,2048,3072 --xflags 0 --ring 2x0,3x0 --reps 50 > $O/sweep_c2.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads c2 --tiles 0 --xflags 33554432,262144 --reps 50 >> $O/sweep_c2.jsonl 2>&1
timeout 300 python tools/sweep.py --workloads ss:boneS10,ss:Cube_Coup_dt6,ss:Bump_2911,ss:RM07R --tiles 0 --xflags 0,33554432,262144 --ring 2x0 --reps 30 > $O/sweep_ss.jsonl 2>&1
python - <<'PY'
import json
for f in ("sweep_c5s", "sweep_c2", "sweep_ss"):
    for l in open(f"gpurun_out/r2c7/{f}.jsonl"):
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
echo "== parity"
timeout 900 python -m pytest tests/test_spmv_gpu.py tests/test_analysis_gpu.py tests/test_fused_halo_gpu.py tests/test_selector_study_gpu.py -q -m gpu --timeout 300 -k "not full_size" > $O/pytest_quick.log 2>&1; echo "rc=$?" >> $O/pytest_quick.log; tail -8 $O/pytest_quick.log
cp gpurun_out/selector_study*.json $O/ 2>/dev/null
ls -la $O
