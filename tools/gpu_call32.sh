#!/bin/bash
# Round 2, GPU call 32 (one B200): ring kernel with line-aligned row dealing in pushed row blocks -- halo / all-gather
# emulation tests, ring tests, quick bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c32
mkdir -p $O
timeout 600 python -m pytest tests/test_fused_halo_gpu.py tests/test_spmv_gpu.py -q -m gpu --timeout 300 -k "not full_size" > $O/pytest.log 2>&1; tail -4 $O/pytest.log
timeout 600 python bench.py --quick > $O/bench_quick.json 2> $O/err.log; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2c32/bench_quick.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("verified"))
PY
