set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_default.json 2> gpurun_out/r02f_default.err; cat gpurun_out/r02f_default.json | cut -c1-200; tail -3 gpurun_out/r02f_default.err
python bench.py > gpurun_out/r02f_default_200.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_launches_default.csv python bench.py --steps 4 --warmup 3 --prewarm 0 --no-cpu --no-parity --conv-interval 0 --no-strict > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_fast -s 6 -c 1 -o gpurun_out/r02f_prof_box256 python bench.py --steps 8 --warmup 3 --prewarm 0 --no-cpu --no-e2e --no-parity --conv-interval 0 --no-strict > /dev/null 2>&1
ncu --set full --clock-control none -k regex:k_step_fast -s 6 -c 1 -o gpurun_out/r02f_prof_box256_strict python bench.py --arithmetic strict --steps 8 --warmup 3 --prewarm 0 --no-cpu --no-e2e --no-parity --conv-interval 0 > /dev/null 2>&1
