set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_ref.err
python bench.py --steps 20 --warmup 5 --verbose-other > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1_other.json
cut -c1-900 gpurun_out/r02_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu.log 2>&1
ncu --set full --clock-control none -k regex:"disc_vail2_kernel|disc_vail3_kernel" -c 1 -o gpurun_out/r02_disc_vail2 -f python tools/bench_disc.py --steps 2 --warmup 2 > /dev/null 2>&1
OM_DISC_VAIL2=3 ncu --set full --clock-control none -k regex:"disc_vail3_kernel" -c 1 -o gpurun_out/r02_disc_vail3 -f python tools/bench_disc.py --steps 2 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"a3_feat_kernel|a3_walk_kernel|a3_post_kernel" -s 6 -c 3 -o gpurun_out/r02_a3_replay -f python tools/bench_a3.py --steps 3 --warmup 1 > /dev/null 2>&1
