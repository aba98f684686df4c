# H1 playback parity tests + the headline step timing
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d.get('sharded_1m'), d['e2e']['ms_per_step_runs'])"
