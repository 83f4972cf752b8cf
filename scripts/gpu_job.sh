set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_default.json 2> gpurun_out/r02f_default.err; cat gpurun_out/r02f_default.json | cut -c1-300
python bench.py > gpurun_out/r02f_default_200.json 2>/dev/null
B="python bench.py --no-cpu --no-e2e --no-parity --conv-interval 0 --steps 100 --warmup 20"
$B --lattice D3Q27 > gpurun_out/r02f_box256_q27.json 2>/dev/null
$B --precision fp32 > gpurun_out/r02f_box256_fp32.json 2>/dev/null
$B --lattice D2Q9 --size 4096 > gpurun_out/r02f_box4096_q9.json 2>/dev/null
$B --arithmetic strict > gpurun_out/r02f_box256_strict.json 2>/dev/null
$B --workload sphere --size 256 --steps 50 --warmup 10 > gpurun_out/r02f_sphere256.json 2>/dev/null
$B --workload step --size 256 --steps 50 --warmup 10 > gpurun_out/r02f_step256.json 2>/dev/null
$B --size 512 --steps 30 --warmup 5 > gpurun_out/r02f_box512.json 2>gpurun_out/r02f_box512.err; tail -2 gpurun_out/r02f_box512.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches_default.csv python bench.py --steps 4 --warmup 3 --prewarm 0 --no-cpu --no-parity --conv-interval 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 6 -c 1 -o gpurun_out/r02f_prof_box256 python bench.py --steps 8 --warmup 3 --prewarm 0 --no-cpu --no-e2e --no-parity --conv-interval 0 > /dev/null 2>&1
