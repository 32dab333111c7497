SH="6144x4096 4096x4096 28672x4096 4096x14336"
run() { timeout 200 python scripts/kbench.py $SH 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['us_per_gemv'])"; }
TG_W4_KERNEL=2 run "mma.sync          "
for c in 1 2; do for s in 0 1; do TG_W4_KERNEL=1 TG_TC_CTAS=$c TG_TC_SPLIT=$s run "tc ctas=$c split=$s"; done; done
TG_W4_KERNEL=0 run "auto              "
