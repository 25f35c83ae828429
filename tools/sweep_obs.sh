# Step-kernel time against observations per group at C3's other sizes: bash tools/sweep_obs.sh [R ...]
for o in ${@:-16 96 112 160 208 224 336 512 800}; do
python bench.py --steps 3 --warmup 3 --iters-per-step 20 --no-cpu-baseline --obs $o 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('obs', $o, 'step_kernel_ms', round(d['kernel_ms']['step_kernel_avg'],4), d['roofline']['bound'], 'frac', round(d['roofline']['frac'],3))"
done
