set -x
timeout 900 python -m pytest tests/test_gpu_operator.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_r18.log 2>&1
tail -25 gpurun_out/pytest_r18.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu > gpurun_out/p_f32_b1.json 2>> gpurun_out/p_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu > gpurun_out/p_f32_b16.json 2>> gpurun_out/p_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu > gpurun_out/p_f32_b128.json 2>> gpurun_out/p_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu > gpurun_out/p_f32_b256.json 2>> gpurun_out/p_err.log
tail -3 gpurun_out/p_err.log
