timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --verbose-other 2> gpurun_out/bench_other.json | cut -c1-2500
