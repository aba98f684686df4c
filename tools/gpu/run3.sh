python -m pytest tests/test_gpu_a3.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python tools/bench_a3.py --steps 30 | grep -o '"task_kernel_ms": [0-9.]*\|"frac": [0-9.]*'
python tools/bench_a3.py --steps 5 --envs 262144 | grep -o '"task_kernel_ms": [0-9.]*\|"frac": [0-9.]*'
python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | grep -o '"e2e": {[^}]*}\|"ms_per_step": [0-9.]*'
python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | grep -o '"e2e": {[^}]*}\|"ms_per_step": [0-9.]*'
