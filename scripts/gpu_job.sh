set -x
timeout 900 python -m pytest tests/test_zzz_distributed_gpu.py tests/test_distributed.py -m gpu -q 2>&1 | tee gpurun_out/r02_pytest_nccl_p2p_2gpu.log | tail -6
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 10 --no-cpu --no-e2e --conv-interval 0"
$T > gpurun_out/r02s_n2_p2p.json 2> gpurun_out/r02s_n2_p2p.err; cat gpurun_out/r02s_n2_p2p.json; tail -3 gpurun_out/r02s_n2_p2p.err
$T --halo nccl > gpurun_out/r02s_n2_nccl.json 2> /dev/null; cat gpurun_out/r02s_n2_nccl.json
