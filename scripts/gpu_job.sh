set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 50 --warmup 10 --no-cpu"
$T > gpurun_out/r02s_n8.json 2> gpurun_out/r02s_n8.err; cat gpurun_out/r02s_n8.json; tail -2 gpurun_out/r02s_n8.err
LBM_B200_NO_OVERLAP=1 $T --no-e2e --no-parity --conv-interval 0 > gpurun_out/r02s_n8_noovl.json 2> /dev/null; cat gpurun_out/r02s_n8_noovl.json
