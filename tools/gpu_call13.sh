#!/bin/bash
# Round 2, GPU call 13 (one B200): ncu of the MIXED kernel on the skewed stand-in (the weak spot of the selector study)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c13
mkdir -p $O
timeout 300 python tools/sweep.py --workloads ss:vas_stokes_2M,ss:Ga41As41H72 --tiles 0,1024,2048,4096 --xflags 0,1048576,2097152,64 --reps 30 > $O/sweep_skewed.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2c13/sweep_skewed.jsonl"):
    try:
        d = json.loads(l)
        print(d["workload"], d["tile"], d["flags"], d["ms"], d["gbs"], d.get("kinds"), d["smem"]) if "ms" in d else print(d)
    except Exception:
        print("??", l[:160])
PY
timeout 300 python tools/profile_one.py ss:vas_stokes_2M > $O/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o $O/ncu_vas_mixed python tools/profile_one.py ss:vas_stokes_2M > $O/ncu.log 2>&1
cat $O/plain.log; tail -2 $O/ncu.log
