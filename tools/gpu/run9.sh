set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 180 -x > gpurun_out/pytest_tc5.log 2>&1
tail -5 gpurun_out/pytest_tc5.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 128 --no-cpu > gpurun_out/z_i8_b128.json 2>> gpurun_out/z_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 256 --no-cpu > gpurun_out/z_i8_b256_cta2.json 2>> gpurun_out/z_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu > gpurun_out/z_i8_b1024_cta2.json 2>> gpurun_out/z_err.log
ncu --set full --clock-control none --import-source on -k regex:scan_i8_tc2 -s 3 -c 1 -o gpurun_out/prof_i8tc2_r01c python bench.py --dtype i8 --batch 256 --rows 2000000 --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_i8tc2c.log 2>&1
tail -5 gpurun_out/z_err.log
