set -x
./tools/microbench/tma_stream 8000000 > gpurun_out/tma_stream.txt 2>&1
cat gpurun_out/tma_stream.txt
timeout 900 python -m pytest tests/test_gpu_tc_f32.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider --timeout 180 > gpurun_out/pytest_f32s.log 2>&1
tail -15 gpurun_out/pytest_f32s.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu > gpurun_out/w_f32_b256_shadow.json 2>> gpurun_out/w_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 128 --no-cpu > gpurun_out/w_f32_b128_shadow.json 2>> gpurun_out/w_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu --opt use_shadow=0 > gpurun_out/w_f32_b256_tf32.json 2>> gpurun_out/w_err.log
tail -5 gpurun_out/w_err.log
