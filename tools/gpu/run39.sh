set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 -x > $O/pytest_r39.log 2>&1
tail -3 $O/pytest_r39.log
B="timeout 300 python bench.py --no-cpu --steps 30"
$B > $O/d_f32_b256.json 2> $O/d.err
$B --batch 1 > $O/d_f32_b1.json 2>> $O/d.err
$B --batch 16 > $O/d_f32_b16.json 2>> $O/d.err
$B --dtype i8 --batch 1024 > $O/d_i8_b1024.json 2>> $O/d.err
$B --dtype i8 --batch 256 > $O/d_i8_b256.json 2>> $O/d.err
$B --rows 1250000 > $O/d_f32_b256_shard8.json 2>> $O/d.err
tail -n 3 $O/d.err
python tools/summarize.py $O/d_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*" | paste - - - - -
