set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02s_default.json 2> gpurun_out/r02s_default.err; cat gpurun_out/r02s_default.json; tail -3 gpurun_out/r02s_default.err
B="python bench.py --no-cpu --no-e2e --no-parity --conv-interval 0 --steps 100 --warmup 20"
$B --workload sphere --size 256 --steps 50 --warmup 10 > gpurun_out/r02s_sphere256.json 2>/dev/null
$B --workload step --size 256 --steps 50 --warmup 10 > gpurun_out/r02s_step256.json 2>/dev/null
python -m pytest tests/test_gpu_parity.py tests/test_zz_baseline_configs_gpu.py tests/test_host_run_gpu.py tests/test_validation_gpu.py -m gpu -x -q > gpurun_out/r02s_pytest_subset.log 2>&1; tail -3 gpurun_out/r02s_pytest_subset.log
