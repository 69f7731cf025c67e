set -x
timeout 900 python -m pytest tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 180 -x > gpurun_out/pytest_tc4.log 2>&1
tail -30 gpurun_out/pytest_tc4.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 256 --no-cpu > gpurun_out/y_i8_b256_cta2.json 2>> gpurun_out/y_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 256 --no-cpu --opt tc_cta2=0 > gpurun_out/y_i8_b256_cta1.json 2>> gpurun_out/y_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu > gpurun_out/y_i8_b1024_cta2.json 2>> gpurun_out/y_err.log
tail -5 gpurun_out/y_err.log
