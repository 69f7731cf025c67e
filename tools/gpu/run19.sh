set -x
K='regex:scan_|select_kernel|rescore_kernel|finalize_kernel|prep_queries|reset_'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 3000 --csv --log-file gpurun_out/r01_launches_f32_b256.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_ll_f32.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 3000 --csv --log-file gpurun_out/r01_launches_i8_b1024.csv python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/ncu_ll_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_float_tc2 -c 12 -o gpurun_out/r01_prof_f32_b256 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 -c 12 -o gpurun_out/r01_prof_i8_b1024 python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/ncu_full_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_float_tc_kernel -c 3 -o gpurun_out/r01_prof_f32_b1 python bench.py --steps 2 --warmup 3 --no-cpu --batch 1 > gpurun_out/ncu_full_f32b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc_kernel -c 3 -o gpurun_out/r01_prof_i8_b1 python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 1 > gpurun_out/ncu_full_i8b1.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches*
