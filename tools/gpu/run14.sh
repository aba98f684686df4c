timeout 600 python -m pytest tests/test_gpu_a3.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
bash tools/gpu/run10.sh
bash tools/gpu/run13.sh
