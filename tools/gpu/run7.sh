python tools/bench_a3.py --steps 10 --warmup 3 | grep -o '"task_kernel_ms": [0-9.]*\|"host_enqueue_ms_per_step": [0-9.]*'
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --verbose-other 2>&1 >/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    for k in ('a3_ppo_rollout_16384x64','a3_ppo_rollout_262144x64'):
        print(k, d[k]['task_kernel_ms'], d[k]['host_enqueue_ms_per_step'], d[k]['ms_per_step'])"
