python -m pytest tests/test_gpu_a3.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -8
python tools/bench_a3.py --steps 20; python tools/bench_a3.py --steps 20 --separate-returns | cut -c1-330
python tools/bench_a3.py --steps 5 --envs 262144 | cut -c1-330
