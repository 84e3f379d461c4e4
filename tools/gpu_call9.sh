#!/bin/bash
# Round 2, GPU call 9 (N GPUs of one box): bench.py under torchrun -- fused single-launch loop + graph, NCCL modes
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=${1:-2}
O=gpurun_out/r2c9
mkdir -p $O
nvidia-smi topo -m > $O/topo_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "rc=$?"
tail -5 $O/bench_n$N.err
python - <<PY
import json
d = json.loads(open("$O/bench_n$N.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("verified"), d["config"]["workload"][:40])
for k, v in (d.get("iterated") or {}).items():
    print(" iter", k, v.get("ms_per_iter"), v.get("x_checksum_first_16th"), v.get("launches_per_iteration"), v.get("error"))
for k, v in (d.get("other_configs") or {}).items():
    print(" other", k, v.get("ms"), (v.get("roofline") or {}).get("frac"), v.get("verified"))
print(" e2e", d.get("e2e"))
PY
