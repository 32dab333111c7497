# A/B timing of library build variants (any4_b200/lib_<v>): VARIANTS="a b" SHAPES="4096 8192" MS="1 16"
for v in ${VARIANTS:-main}; do
  if [ "$v" = main ]; then unset ANY4_B200_LIB_DIR; else export ANY4_B200_LIB_DIR=$PWD/any4_b200/lib_$v; fi
  for m in ${MS:-1}; do
    TG_W4_KERNEL=${KERNEL:-1} KB_M=$m timeout 300 python scripts/kbench.py ${SHAPES:-4096 8192 11008} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('variant=$v m=$m', d['us_per_gemv'], all(d['bit_equal_to_plain_launches'].values()))"
  done
done
