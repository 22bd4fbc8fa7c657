for pend in 1 0; do
SOC_PEND=$pend python bench.py --steps 3 --warmup 3 --no-cpu --kernel-times 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('pend', $pend, d['value'], d['cell_steps_per_s'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
