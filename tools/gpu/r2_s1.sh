set -x
nvidia-smi -L
export PKV_SESSION=r2s1
timeout 900 python -m pytest tests/test_gpu_live.py -q --tb=short -p no:cacheprovider --timeout 300 -x > gpurun_out/r2s1_live.log 2>&1
tail -25 gpurun_out/r2s1_live.log
timeout 1200 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 600 --deselect tests/test_gpu_live.py > gpurun_out/r2s1_pytest.log 2>&1
tail -15 gpurun_out/r2s1_pytest.log
timeout 400 python bench.py --no-cpu --steps 20 > gpurun_out/r2s1_f32_b256.json 2> gpurun_out/r2s1_f32_b256.err; tail -3 gpurun_out/r2s1_f32_b256.err
timeout 300 python bench.py --no-cpu --steps 20 --rows 1000000 > gpurun_out/r2s1_f32_b256_1M.json 2> gpurun_out/r2s1_1M.err; tail -3 gpurun_out/r2s1_1M.err
timeout 400 python bench.py --no-cpu --steps 10 --dtype i8 --batch 1024 > gpurun_out/r2s1_i8_b1024.json 2> gpurun_out/r2s1_i8.err; tail -3 gpurun_out/r2s1_i8.err
timeout 300 python bench.py --no-cpu --steps 20 --opt live=0 > gpurun_out/r2s1_f32_b256_chunked.json 2>> gpurun_out/r2s1_f32_b256.err
cat gpurun_out/r2s1_*.json | cut -c1-900
