set -x
timeout 300 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu > gpurun_out/r_f32_b1_tc.json 2>> gpurun_out/r_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu > gpurun_out/r_f32_b16_tc.json 2>> gpurun_out/r_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 1 --no-cpu > gpurun_out/r_i8_b1_simt.json 2>> gpurun_out/r_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 1 --no-cpu --opt tc_min_queries=1 > gpurun_out/r_i8_b1_tc.json 2>> gpurun_out/r_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 8 --no-cpu > gpurun_out/r_i8_b8_simt.json 2>> gpurun_out/r_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 8 --no-cpu --opt tc_min_queries=1 > gpurun_out/r_i8_b8_tc.json 2>> gpurun_out/r_err.log
# launch lists (every launch with its device time; shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pkv -c 400 --csv --log-file gpurun_out/launches_f32_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ll_f32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pkv -c 600 --csv --log-file gpurun_out/launches_i8_b1024.csv python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/ncu_ll_i8.log 2>&1
# full captures of the dominant kernels (last = biggest chunk)
ncu --set full --clock-control none --import-source on -k regex:scan_float_tc2 -s 20 -c 2 -o gpurun_out/prof_r01_f32_b256_tc2 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 -s 80 -c 2 -o gpurun_out/prof_r01_i8_b1024_tc2 python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/ncu_full_i8.log 2>&1
tail -3 gpurun_out/r_err.log gpurun_out/ncu_full_f32.log gpurun_out/ncu_full_i8.log
ls -la gpurun_out | tail -12
