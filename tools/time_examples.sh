# Wall time of the reference's two example workloads (BASELINE configs 1 and 2): the GPU engine
# (examples/*.py: process start, CUDA initialisation, start state, sampling, files, diagnoseSamples)
# and, with "ref" as first argument, the unmodified reference on the same box beside it.
# usage (GPU box): bash tools/time_examples.sh [ref]
mode=$1
cd /tmp
python -c "import torch; torch.zeros(1).cuda()"      # page the image in
for ex in "distribution partial" "regression partial" "regression none" "regression complete"; do
  set -- $ex
  t0=$(date +%s.%N)
  python $GRAFT_REPO_ROOT/examples/$1.py $2 > /tmp/out_$1_$2.txt 2> /tmp/err_$1_$2.txt
  rc=$?
  t1=$(date +%s.%N)
  echo "GPU_EXAMPLE $1 $2 rc=$rc wall $(python -c "print('%.2f' % ($t1 - $t0))") s, $(grep -c . /tmp/out_$1_$2.txt) lines of output"
done
tail -8 /tmp/out_regression_partial.txt
if [ "$mode" = "ref" ]; then
  for ex in "distribution partial" "regression partial" "regression none" "regression complete"; do
    set -- $ex
    python $GRAFT_REPO_ROOT/baseline/run_reference_examples.py $1 $2 2>/dev/null | grep REFERENCE_EXAMPLE
  done
fi
