# Round-2 final evidence run (1 GPU): smoke, tests, both bench arms, launch list of the bench command, full captures of the top kernels.
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu.log; cat gpurun_out/r02_pytest_gpu.log
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_ref.err
python bench.py --steps 20 --warmup 5 --verbose-other > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1_other.json
cut -c1-1200 gpurun_out/r02_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:play_h1_tp_kernel -s 4 -c 1 -o gpurun_out/r02_play_h1_tp -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"play_moments_kernel|play_snapshot_kernel|affine_scan_kernel|normalize_kernel|moments_kernel" -s 20 -c 5 -o gpurun_out/r02_step_tail -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"a3_feat_kernel|a3_walk_kernel|a3_post_kernel" -s 6 -c 3 -o gpurun_out/r02_a3_replay -f python tools/bench_a3.py --steps 3 --warmup 1 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,smsp__inst_executed.sum --cache-control none --clock-control none -k regex:"a3_|affine_scan" -s 12 -c 12 --csv --log-file gpurun_out/r02_a3_inflow.csv python tools/bench_a3.py --steps 6 --warmup 3 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"disc_vail4_kernel|disc_reward_pg2_kernel" -s 3 -c 2 -o gpurun_out/r02_disc -f python tools/bench_disc.py --steps 2 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"disc_vail4_kernel" -s 2 -c 1 -o gpurun_out/r02_disc_vail4_1m -f python tools/bench_disc.py --envs 1048576 --steps 2 --warmup 2 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:"h1_live_step_kernel" -s 6 -c 1 -o gpurun_out/r02_h1_live_step -f python tools/bench_h1_step.py --steps 5 > /dev/null 2>&1
tools/micro/umma_rate > gpurun_out/r02_umma_rate.txt 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
