# A3 parity tests + the rollout timing at both sizes
timeout 600 python -m pytest tests/test_gpu_a3.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for n in 16384 262144; do timeout 200 python tools/bench_a3.py --envs $n --steps 10 --warmup 3 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('task_kernel_ms','ms_per_step','eager_task_kernel_ms')}, d['roofline']['frac'])"; done
