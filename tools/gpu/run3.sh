python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02c_pytest.log
cat gpurun_out/r02c_pytest.log
python tools/bench_a3.py --steps 20 > gpurun_out/r02c_a3.json 2>&1; cat gpurun_out/r02c_a3.json
