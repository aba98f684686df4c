# H1 parity tests + single-step timings at both shard sizes + the headline step timing
timeout 900 python -m pytest tests/test_gpu_env.py tests/test_gpu_h1.py tests/test_gpu_fk.py tests/test_gpu_abi_ld.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for n in 1048576 131072; do timeout 600 python tools/bench_h1_step.py --envs $n --steps 20 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['workload'][:40], d['h1_step_kernel_ms'], d['roofline']['frac'], d['live_step_ms'], d['live_step_frac'])"; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step_runs'])"
