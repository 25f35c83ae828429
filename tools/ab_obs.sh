for R in 200 160; do
for lib in "" "$@"; do
  if [ -z "$lib" ]; then unset MCMCN_LIB; else export MCMCN_LIB=$PWD/$lib; fi
  timeout 200 python bench.py --obs $R --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($R, '${lib:-default}', d['value'], d['kernel_ms']['step_kernel_avg'])"
done; done
