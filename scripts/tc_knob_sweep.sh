mkdir -p gpurun_out/r2
export TG_W4_KERNEL=1
for ctas in 1 2; do for split in 0 1; do
  TG_TC_CTAS=$ctas TG_TC_SPLIT=$split timeout 200 python scripts/kbench.py 4096 8192 11008 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ctas=$ctas split=$split', d['us_per_gemv'], d['GBps'])"
done; done
