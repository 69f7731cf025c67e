set -x
timeout 600 python -m pytest tests/test_gpu_tc_f32.py tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 180 -x > gpurun_out/pytest_bs.log 2>&1
tail -4 gpurun_out/pytest_bs.log
for g in 0 600 400; do
timeout 300 python bench.py --steps 5 --warmup 3 --batch 256 --no-cpu --opt chunk_growth_x100=$g > gpurun_out/v_f32_b256_g$g.json 2>> gpurun_out/v_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 256 --no-cpu --opt chunk_growth_x100=$g > gpurun_out/v_i8_b256_g$g.json 2>> gpurun_out/v_err.log
timeout 300 python bench.py --steps 5 --warmup 3 --dtype i8 --batch 1024 --no-cpu --opt chunk_growth_x100=$g > gpurun_out/v_i8_b1024_g$g.json 2>> gpurun_out/v_err.log
done
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace2_f32_b256.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --dtype i8 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace2_i8_b256.log
tail -4 gpurun_out/trace2_f32_b256.log gpurun_out/trace2_i8_b256.log
tail -3 gpurun_out/v_err.log
