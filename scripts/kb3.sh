# quick timing of the three headline shapes (both kernels) + correctness of the big-shape goldens
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_big_shapes_gpu.py tests/test_gemm_gpu.py tests/test_decode_gpu.py -q -x --timeout 300 2>&1 | tail -3
for kk in ${KERNELS:-0 1}; do for m in ${MS:-1}; do
  TG_W4_KERNEL=$kk KB_M=$m timeout 200 python scripts/kbench.py 4096 8192 11008 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('kernel=$kk m=$m', d['us_per_gemv'], d['GBps'], d['bit_equal_to_plain_launches'])"
done; done
