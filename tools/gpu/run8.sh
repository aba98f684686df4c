python tools/bench_a3.py --steps 10 --warmup 3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('task_kernel_ms','ms_per_step','host_enqueue_ms_per_step','eager_task_kernel_ms','timing','gpu_launches')}, d['roofline']['frac'])"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose-other 2>&1 >/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    for k in ('a3_ppo_rollout_16384x64','a3_ppo_rollout_262144x64'):
        print(k, {j:d[k][j] for j in ('task_kernel_ms','ms_per_step','host_enqueue_ms_per_step','eager_task_kernel_ms','timing')}, d[k]['roofline']['frac'])"
