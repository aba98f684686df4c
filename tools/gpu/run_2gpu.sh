# 2-GPU: torchrun-gated mailbox test (in-process on both ranks), the self-launching variant, check_mailbox latency line, bench at N=2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 -m pytest tests/test_gpu_mailbox_multi.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02b_mailbox_torchrun.log
cat gpurun_out/r02b_mailbox_torchrun.log
python -m pytest tests/test_gpu_mailbox_multi.py tests/test_gpu_mailbox.py -m gpu -x -q 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 tools/check_mailbox.py 2>&1 | tail -1 | tee gpurun_out/r02b_check_mailbox.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_n2.json 2> gpurun_out/r02b_bench_n2.err
tail -c 300 gpurun_out/r02b_bench_n2.err; cut -c1-2200 gpurun_out/r02b_bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 | cut -c1-300
