set -x
timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/pytest_r15.log 2>&1
tail -8 gpurun_out/pytest_r15.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 256 --no-cpu > gpurun_out/s_i8_b256.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --dtype i8 --batch 1024 --no-cpu > gpurun_out/s_i8_b1024.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu > gpurun_out/s_f32_b256.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 256 --no-cpu --opt optimistic=0 > gpurun_out/s_f32_b256_careful.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu > gpurun_out/s_f32_b128.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 1 --no-cpu > gpurun_out/s_f32_b1.json 2>> gpurun_out/s_err.log
timeout 300 python bench.py --steps 10 --warmup 3 --batch 16 --no-cpu > gpurun_out/s_f32_b16.json 2>> gpurun_out/s_err.log
tail -3 gpurun_out/s_err.log
