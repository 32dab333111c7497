#!/bin/bash
# One GPU-box visit: smoke, parity tests, reference goldens, bench, ncu.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
if [ "${SKIP_TESTS:-0}" != "1" ]; then
if [ "${SANITIZE:-0}" = "1" ]; then echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gemm_gpu.py -q -x -k "n64-k4096-g128-ik4 and any4r-bf16" > gpurun_out/sanitizer.log 2>&1; tail -30 gpurun_out/sanitizer.log; fi
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi
if [ "${GOLDEN:-0}" = "1" ]; then
echo "== reference goldens"; timeout 900 python oracle/ref_runner.py golden gpurun_out/golden_gpu.npz > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -5 gpurun_out/golden.log
fi
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "${REFBENCH:-0}" = "1" ]; then
for s in 4096 8192 11008; do timeout 300 python oracle/ref_runner.py bench $s $s 10 >> gpurun_out/ref_gpu_bench.jsonl 2>> gpurun_out/ref_gpu_bench.err; done; cat gpurun_out/ref_gpu_bench.jsonl
fi
if [ "${NCU:-0}" = "1" ]; then
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemv_w4_b|gemm_frag|_kernel" -s 300 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-sweep --no-llama > gpurun_out/ncu_bench.log 2>&1; echo "ncu1 rc=$?"
echo "== ncu full 8192"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_b -s 12 -c 2 -o gpurun_out/prof8192 python bench.py --steps 2 --warmup 3 --no-graph --profile-shape 8192 > gpurun_out/ncu_full8192.log 2>&1; echo "ncu3 rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_b -s 60 -c 3 -o gpurun_out/prof python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-sweep > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
fi
echo done
