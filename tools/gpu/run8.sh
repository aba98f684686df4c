# 8-GPU: D2H ceiling probe, then the bench at N=8 (mailbox) for the record
nvidia-smi topo -m > gpurun_out/r02d_topo.txt 2>&1; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/r02d_lscpu.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/probe_d2h.py 2> gpurun_out/r02d_probe.err | tail -1 > gpurun_out/r02d_probe_d2h.json
cat gpurun_out/r02d_probe_d2h.json; tail -3 gpurun_out/r02d_probe.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_n8.json 2> gpurun_out/r02d_bench_n8.err
tail -c 300 gpurun_out/r02d_bench_n8.err; cut -c1-2500 gpurun_out/r02d_bench_n8.json
head -30 gpurun_out/r02d_topo.txt; cat gpurun_out/r02d_lscpu.txt
