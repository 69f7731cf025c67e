set -x
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 -s 3 -c 1 -o gpurun_out/prof_i8tc2_r01b python bench.py --dtype i8 --batch 256 --rows 2000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_i8tc2.log 2>&1
tail -3 gpurun_out/ncu_i8tc2.log
ncu --set full --clock-control none --import-source on -k regex:scan_f32_tc -s 3 -c 1 -o gpurun_out/prof_f32tc_r01b python bench.py --batch 256 --rows 2000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_f32tc.log 2>&1
tail -3 gpurun_out/ncu_f32tc.log
