set -x
timeout 1500 python -m pytest tests/test_gpu_tc_f32.py tests/test_gpu_parity.py -q --tb=short -p no:cacheprovider --timeout 600 -x > gpurun_out/pytest_r32.log 2>&1
tail -5 gpurun_out/pytest_r32.log
python __graft_entry__.py smoke 2>&1 | tail -2
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/a_f32_b256.json 2> gpurun_out/a.err
$B --batch 1024 > gpurun_out/a_f32_b1024.json 2>> gpurun_out/a.err
$B --batch 16 > gpurun_out/a_f32_b16.json 2>> gpurun_out/a.err
$B --rows 1250000 > gpurun_out/a_f32_b256_shard8.json 2>> gpurun_out/a.err
$B --rows 1250000 --dtype i8 --batch 1024 > gpurun_out/a_i8_b1024_shard8.json 2>> gpurun_out/a.err
$B --dtype i8 --batch 1024 > gpurun_out/a_i8_b1024.json 2>> gpurun_out/a.err
$B --bitmap-density 0.1 > gpurun_out/a_f32_b256_bm10.json 2>> gpurun_out/a.err
$B --bitmap-density 0.5 --dtype i8 --batch 1024 > gpurun_out/a_i8_b1024_bm50.json 2>> gpurun_out/a.err
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 10 > gpurun_out/a_f16_b4096_shard.json 2>> gpurun_out/a.err
tail -n 5 gpurun_out/a.err
python tools/summarize.py gpurun_out/a_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|launches [0-9.]*\|rescans [0-9]*" | paste - - - - - -
