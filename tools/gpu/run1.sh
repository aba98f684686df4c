python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02a_pytest.log
cat gpurun_out/r02a_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/r02a_launches_bench.csv | head -30
