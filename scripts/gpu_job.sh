N=$1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu --no-e2e --conv-interval 0 --workload step --size 256 > gpurun_out/r02f_step256_n$N.json 2> gpurun_out/r02f_step256_n$N.err; cat gpurun_out/r02f_step256_n$N.json | cut -c1-200; tail -2 gpurun_out/r02f_step256_n$N.err
