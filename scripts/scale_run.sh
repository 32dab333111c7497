#!/bin/bash
# bench.py at N GPUs (the driver's launch line) -> gpurun_out/r2/bench_${N}gpu.json
N=$1; shift
mkdir -p gpurun_out/r2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/r2/bench_${N}gpu.json 2> gpurun_out/r2/bench_${N}gpu.err
echo "rc=$?"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2/bench_${N}gpu.err | tail -3
python - <<PY
import json
d = json.loads(open("gpurun_out/r2/bench_${N}gpu.json").read().strip().splitlines()[-1])
print(json.dumps({k: d.get(k) for k in ("n_gpus", "value", "parity_checked")}), d["roofline"]["us_per_launch"], d["config"]["other_shapes"])
print(json.dumps(d["config"].get("llama_decode")))
PY
