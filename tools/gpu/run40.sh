set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc_f32.py tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 600 -x -k "f32 or img8 or shapes or float" > $O/pytest_r40.log 2>&1
tail -3 $O/pytest_r40.log
B="timeout 300 python bench.py --no-cpu --steps 30"
$B > $O/e_f32_b256.json 2> $O/e.err
$B --batch 1024 > $O/e_f32_b1024.json 2>> $O/e.err
$B --batch 16 > $O/e_f32_b16.json 2>> $O/e.err
$B --batch 128 > $O/e_f32_b128.json 2>> $O/e.err
$B --rows 1250000 > $O/e_f32_b256_shard8.json 2>> $O/e.err
tail -n 3 $O/e.err
python tools/summarize.py $O/e_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*" | paste - - - - -
