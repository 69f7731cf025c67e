set -x
timeout 600 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_operator.py -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/r2s14_tests.log 2>&1
tail -40 gpurun_out/r2s14_tests.log
