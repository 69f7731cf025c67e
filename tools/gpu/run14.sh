set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 180 -x > gpurun_out/pytest_r14.log 2>&1
tail -6 gpurun_out/pytest_r14.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 128 --no-cpu > gpurun_out/t_i8_b128.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 256 --no-cpu > gpurun_out/t_i8_b256.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 1024 --no-cpu > gpurun_out/t_i8_b1024.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu > gpurun_out/t_f32_b256_pair.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu --opt tc_cta2=0 > gpurun_out/t_f32_b256_single.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu > gpurun_out/t_f32_b128.json 2>> gpurun_out/t_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 1024 --no-cpu > gpurun_out/t_f32_b1024.json 2>> gpurun_out/t_err.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace3_f32_b256.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --dtype i8 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace3_i8_b256.log
tail -3 gpurun_out/t_err.log
