timeout 120 python tools/bench_a3.py --envs 262144 --steps 10 --warmup 3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('task_kernel_ms','ms_per_step','eager_task_kernel_ms')}, d['roofline']['frac'])"
timeout 120 python tools/bench_a3.py --envs 1000 --horizon 200 --steps 10 --warmup 3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('task_kernel_ms','ms_per_step','eager_task_kernel_ms')}, d['roofline']['frac'])"
