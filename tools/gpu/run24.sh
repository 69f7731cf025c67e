set -x
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_i8_ts --launch-skip 4 --launch-count 1 -o gpurun_out/r01_prof_i8_ts_g4 -f python bench.py --dtype i8 --batch 1024 --no-cpu --steps 1 --warmup 1 --opt ts_groups=4 > gpurun_out/ncu_ts.log 2>&1
tail -n 3 gpurun_out/ncu_ts.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_float_tc2 --launch-skip 4 --launch-count 1 -o gpurun_out/r01_prof_f32_b256_v2 -f python bench.py --no-cpu --steps 1 --warmup 1 > gpurun_out/ncu_f32.log 2>&1
tail -n 3 gpurun_out/ncu_f32.log
ls -la gpurun_out/*.ncu-rep
