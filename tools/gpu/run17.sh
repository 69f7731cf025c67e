set -x
timeout 600 python -m pytest tests/test_gpu_operator.py tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_r17.log 2>&1
tail -5 gpurun_out/pytest_r17.log
ncu --set full --clock-control none --import-source on -k regex:scan_float_tc2 -s 5 -c 1 -o gpurun_out/prof_r01_f32_b256_main python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_full_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 -s 5 -c 1 -o gpurun_out/prof_r01_i8_b256_main python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 256 > gpurun_out/ncu_full_i8.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc_kernel -s 5 -c 1 -o gpurun_out/prof_r01_i8_b128_main python bench.py --steps 2 --warmup 3 --no-cpu --dtype i8 --batch 128 > gpurun_out/ncu_full_i8b.log 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu > gpurun_out/q_f32_b1.json 2>> gpurun_out/q_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 1 --no-cpu > gpurun_out/q_i8_b1.json 2>> gpurun_out/q_err.log
tail -3 gpurun_out/q_err.log
