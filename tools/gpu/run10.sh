python tools/bench_a3.py --steps 10 --warmup 3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('task_kernel_ms','ms_per_step','eager_task_kernel_ms')}, d['roofline']['frac'])"
