set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_staged.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_staged.log
B="python bench.py --no-cpu --no-e2e --steps 100 --warmup 20"
$B > gpurun_out/r02s_box256.json 2>gpurun_out/r02s_box256.err; cat gpurun_out/r02s_box256.json
LBM_B200_LIB=$PWD/lbm_b200/variants/lib_t256.so $B > gpurun_out/r02s_box256_t256.json 2>/dev/null; cat gpurun_out/r02s_box256_t256.json
LBM_B200_DEBUG_IDENTITY=1 $B > gpurun_out/r02s_box256_identity.json 2>/dev/null; cat gpurun_out/r02s_box256_identity.json
$B --lattice D3Q27 > gpurun_out/r02s_box256_q27.json 2>/dev/null; cat gpurun_out/r02s_box256_q27.json
$B --precision fp32 > gpurun_out/r02s_box256_fp32.json 2>/dev/null; cat gpurun_out/r02s_box256_fp32.json
$B --lattice D2Q9 --size 4096 > gpurun_out/r02s_box4096_q9.json 2>/dev/null; cat gpurun_out/r02s_box4096_q9.json
$B --arithmetic strict > gpurun_out/r02s_box256_strict.json 2>/dev/null; cat gpurun_out/r02s_box256_strict.json
$B --workload sphere --size 256 --steps 50 --warmup 10 > gpurun_out/r02s_sphere256.json 2>/dev/null; cat gpurun_out/r02s_sphere256.json
$B --workload step --size 256 --steps 50 --warmup 10 > gpurun_out/r02s_step256.json 2>/dev/null; cat gpurun_out/r02s_step256.json
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 6 -c 1 -o gpurun_out/r02s_prof python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e > /dev/null 2>&1
ls gpurun_out | head -50
