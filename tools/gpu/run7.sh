python -m pytest tests/test_gpu_learner.py tests/test_gpu_a3.py tests/test_gpu_h1.py tests/test_gpu_env.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*'; done
python tools/bench_a3.py --steps 30 | grep -o '"task_kernel_ms": [0-9.]*\|"frac": [0-9.]*'
python tools/bench_h1_step.py --envs 131072 | grep -o '"live_step_ms": [0-9.]*\|"live_step_eager_ms": [0-9.]*\|"live_step_frac": [0-9.]*'
python tools/bench_h1_step.py | grep -o '"live_step_ms": [0-9.]*\|"live_step_eager_ms": [0-9.]*\|"live_step_frac": [0-9.]*'
