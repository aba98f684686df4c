python -m pytest tests/test_gpu_a3.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do python tools/bench_a3.py --steps 30 | grep -o '"task_kernel_ms": [0-9.]*\|"frac": [0-9.]*'; done
python tools/bench_a3.py --steps 5 --envs 262144 | grep -o '"task_kernel_ms": [0-9.]*\|"frac": [0-9.]*'
