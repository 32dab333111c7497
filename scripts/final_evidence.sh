#!/bin/bash
# End-of-round single-GPU evidence: the driver's bench line, the ncu launch list of the same command, the Llama
# layer timeline.  -> gpurun_out/r2/
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2; mkdir -p $OUT
timeout 900 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; echo "bench rc=$?"; tail -2 $OUT/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_reference_arm.json 2>> $OUT/bench_1gpu.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_|gemm_|stage_host|x_permute|add_rmsnorm|rope_attn|silu_mul" -c 400 --csv --log-file $OUT/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-sweep --no-llama > $OUT/ncu_launches.log 2>&1; echo "ncu rc=$?"
python scripts/llama_timeline.py --layers 4 --show 24 > $OUT/llama_layer_timeline.txt 2>&1
timeout 300 python bench_llama.py --impl reference > $OUT/llama_1gpu_reference_kernels.json 2>> $OUT/bench_1gpu.err; echo "llama reference kernels rc=$?"; cat $OUT/llama_1gpu_reference_kernels.json | cut -c1-300
python - <<PY
import json
d = json.loads(open("$OUT/bench_1gpu.json").read())
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "e2e", "gpu_launches")}))
print(json.dumps(d["roofline"]))
print(json.dumps(d["config"]["format_sweep_us"]))
print(json.dumps(d["config"]["llama_decode"]))
print(d["config"]["default_options"], d["config"]["other_shapes"], d["config"]["two_independent_chains(informational)"])
PY
