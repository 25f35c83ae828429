# A/B on BASELINE config 5 of library builds x observations per task (MCMCN_TASK_OBS): bash tools/ab_c5_knobs.sh [lib.so ...]
for lib in "" "$@"; do
  if [ -z "$lib" ]; then unset MCMCN_LIB; else export MCMCN_LIB=$PWD/$lib; fi
  for obs in ${TASK_OBS_LIST:-256 512 1024}; do
    for pooling in partial none; do
      MCMCN_TASK_OBS=$obs timeout 200 python bench.py --workload c5 --pooling $pooling --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('${lib:-default}', 'task_obs=$obs', '$pooling', round(d['value']), round(d['kernel_ms']['step_kernel_avg'], 4))"
    done
  done
done
