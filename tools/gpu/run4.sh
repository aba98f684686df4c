python tools/time_sharded.py 131072 4; python tools/time_sharded.py 131072 2; python tools/time_sharded.py 1048576 2
python -m pytest tests/test_gpu_env.py tests/test_gpu_h1.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-other-configs --no-cpu-baseline 2>/dev/null | cut -c1-1700
