set -x
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc -s 6 -c 2 -o gpurun_out/prof_i8tc_r01a python bench.py --dtype i8 --batch 128 --rows 2000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_i8tc.log 2>&1
tail -5 gpurun_out/ncu_i8tc.log
ls -la gpurun_out/
