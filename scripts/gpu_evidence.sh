#!/bin/bash
# Round evidence on one GPU: the reference's own kernel test-suite against this library, compute-sanitizer over the
# split-k / multi-row-block / graph-replay cases, ncu captures of the kernels besides the headline one.  -> gpurun_out/r2/
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/r2; mkdir -p $OUT
if [ -d oracle/_ref/tests_tinygemm ]; then
  echo "== reference tests/tinygemm/*.py (unmodified copies, git-ignored) against any4_b200"
  (cd oracle/_ref/tests_tinygemm && PYTHONPATH=$PWD/../../.. timeout 1500 python -m pytest -q -p no:cacheprovider . ) > $OUT/reference_testsuite.log 2>&1; echo "rc=$?"; tail -4 $OUT/reference_testsuite.log
fi
if [ -d oracle/_ref/ref_pkg ] && [ "${REF_SELF_TEST:-1}" = "1" ]; then
  echo "== the same suite against the REFERENCE's own extension (which of its tests are flaky on this GPU / cuBLAS?)"
  (cd oracle/_ref/tests_tinygemm && PYTHONPATH=$PWD/../ref_pkg timeout 1500 python -m pytest -q -p no:cacheprovider . ) > $OUT/reference_testsuite_on_reference.log 2>&1; echo "rc=$?"; tail -6 $OUT/reference_testsuite_on_reference.log
fi
echo "== compute-sanitizer memcheck"
SEL='graph_replay or tcgen05_kernel_agrees or (big and n4096-k4096 and (m1 or m16) and right) or (big and n11008 and any4r and m1 and right)'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py tests/test_big_shapes_gpu.py tests/test_decode_gpu.py -q -x --timeout 1400 -k "$SEL" > $OUT/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" $OUT/sanitizer_memcheck.log | tail -3
echo "== compute-sanitizer racecheck"
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 1100 -k "tcgen05_kernel_agrees and any4r" > $OUT/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" $OUT/sanitizer_racecheck.log | tail -3
echo "== ncu: m = 8 / 16 (tensor pipe), A layout, int8 / bf16 streaming kernel, headline mma.sync kernel"
for m in 8 16; do KB_M=$m timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_tc -s 40 -c 1 -f -o $OUT/tc4096_m$m python scripts/kbench.py 4096 > $OUT/ncu_tc4096_m$m.log 2>&1; echo "m$m rc=$?"; done
KB_SIDE=A timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_a -s 40 -c 1 -f -o $OUT/a4096 python scripts/kbench.py 4096 > $OUT/ncu_a4096.log 2>&1; echo "A rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_w4_b_kernel -s 40 -c 2 -f -o $OUT/b4096 python bench.py --steps 2 --warmup 3 --no-graph --profile-shape 4096 > $OUT/ncu_b4096.log 2>&1; echo "B rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_stream -s 6 -c 2 -f -o $OUT/stream python scripts/kbench_generic.py > $OUT/ncu_stream.log 2>&1; echo "stream rc=$?"
echo done
