for o in 4 40 200 800 1600; do
python bench.py --steps 3 --warmup 3 --iters-per-step 20 --no-cpu-baseline --obs $o 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('obs', $o, 'step_kernel_ms', round(d['kernel_ms']['step_kernel_avg'],4), 'achieved TF', round(d['roofline']['achieved'],2))"
done
