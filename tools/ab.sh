# A/B bench of library builds with the same ABI: bash tools/ab.sh [lib.so ...]
for lib in "" "$@"; do
  if [ -z "$lib" ]; then unset MCMCN_LIB; else export MCMCN_LIB=$PWD/$lib; fi
  timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${lib:-default}', d['value'], d['kernel_ms']['step_kernel_avg'], d['kernel_ms']['hyper_kernel_avg'])"
done
