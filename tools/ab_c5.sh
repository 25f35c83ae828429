# A/B of library builds on BASELINE config 5 (Bernoulli-logit): bash tools/ab_c5.sh [lib.so ...]
for lib in "" "$@"; do
  if [ -z "$lib" ]; then unset MCMCN_LIB; else export MCMCN_LIB=$PWD/$lib; fi
  for pooling in partial none; do
  timeout 200 python bench.py --workload c5 --pooling $pooling --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${lib:-default}', '$pooling', d['value'], d['kernel_ms']['step_kernel_avg'])"
  done
done
