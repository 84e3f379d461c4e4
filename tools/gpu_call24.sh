#!/bin/bash
# Round 2, GPU call 24 (one B200): final code -- smoke(), full GPU test suite, full bench line, reference arm
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c24
mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/smoke.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $O/pytest_full.log 2>&1; echo "rc=$?" >> $O/pytest_full.log; tail -5 $O/pytest_full.log
timeout 1200 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
tail -3 $O/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c24/bench_n1.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("traffic"), d.get("verified"))
print(" e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k, v in (d.get("other_configs") or {}).items():
    print(" other", k, v.get("ms"), (v.get("roofline") or {}).get("frac"), v.get("verified"))
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref rc=$?"
cut -c1-300 $O/bench_ref.json
