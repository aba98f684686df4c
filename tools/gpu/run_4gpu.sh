# 4 GPUs: the torchrun-gated mailbox test and bench.py --gpus 4 (weak scaling + the sharded 1M-env rollout)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 -m pytest tests/test_gpu_mailbox_multi.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02e_mailbox_torchrun_n4.log
cat gpurun_out/r02e_mailbox_torchrun_n4.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_n4.json 2> gpurun_out/r02e_bench_n4.err
tail -c 300 gpurun_out/r02e_bench_n4.err; cut -c1-1500 gpurun_out/r02e_bench_n4.json
