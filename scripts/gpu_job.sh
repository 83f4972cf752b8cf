set -x
B="python bench.py --no-cpu --no-e2e --no-parity --conv-interval 0 --steps 100 --warmup 20"
$B --lattice D3Q27 > gpurun_out/r02f_box256_q27.json 2>/dev/null
$B --lattice D2Q9 --size 4096 > gpurun_out/r02f_box4096_q9.json 2>/dev/null
$B --workload sphere --size 256 --steps 50 --warmup 10 > gpurun_out/r02f_sphere256.json 2>/dev/null
$B --lattice D3Q27 --precision fp32 > gpurun_out/r02f_box256_q27_fp32.json 2>/dev/null
$B > gpurun_out/r02f_box256.json 2>/dev/null
python -m pytest tests/test_gpu_parity.py tests/test_zz_baseline_configs_gpu.py tests/test_validation_gpu.py tests/test_zz_multilevel_gpu.py -m gpu -x -q > gpurun_out/r02s_pytest_subset.log 2>&1; tail -3 gpurun_out/r02s_pytest_subset.log
