#!/bin/bash
# N-GPU check of the sharded refinement and the bench's refine / vs_single_gpu block:  bash tests/scale_check.sh 4 [tag]
N=${1:-2}; TAG=${2:-scale}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_refine_check.py 2>&1 | tail -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_n$N.json").read().splitlines() if l.startswith("{")][-1])
for k, v in d["refine"].items():
    print(k, {a: v[a] for a in v if a in ("iters_per_s", "us_per_iter", "ok", "max_rel_diff_vs_single_gpu", "trajectory_abs_diff_mm")})
print("value", d["value"], "e2e", d["e2e"]["value"])
PY
