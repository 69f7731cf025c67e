set -x
timeout 900 python -m pytest tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 180 > gpurun_out/pytest_f32s2.log 2>&1
tail -15 gpurun_out/pytest_f32s2.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace_f32_b256.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --dtype i8 --batch 256 --no-cpu > /dev/null 2> gpurun_out/trace_i8_b256.log
PKV_TRACE=1 timeout 300 python bench.py --steps 2 --warmup 3 --batch 8 --no-cpu > /dev/null 2> gpurun_out/trace_f32_b8.log
tail -8 gpurun_out/trace_f32_b256.log gpurun_out/trace_i8_b256.log gpurun_out/trace_f32_b8.log
