timeout 600 python -m pytest tests/test_gpu_learner.py tests/test_gpu_abi_ld.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs | cut -c1-200
