set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 -x > $O/pytest_r36.log 2>&1
tail -5 $O/pytest_r36.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > $O/c_f32_b256.json 2> $O/c.err
$B --opt ts_acc_buffers=2 > $O/c_f32_b256_nb2.json 2>> $O/c.err
$B --batch 1024 > $O/c_f32_b1024.json 2>> $O/c.err
$B --batch 128 > $O/c_f32_b128.json 2>> $O/c.err
$B --batch 16 > $O/c_f32_b16.json 2>> $O/c.err
$B --batch 16 --opt ts_acc_buffers=2 > $O/c_f32_b16_nb2.json 2>> $O/c.err
$B --dim 512 > $O/c_f32_b256_d512.json 2>> $O/c.err
$B --dtype i8 --batch 1024 > $O/c_i8_b1024.json 2>> $O/c.err
$B --dtype i8 --batch 1024 --opt ts_acc_buffers=2 > $O/c_i8_b1024_nb2.json 2>> $O/c.err
$B --dtype i8 --batch 256 > $O/c_i8_b256.json 2>> $O/c.err
$B --dtype i8 --batch 1024 --dim 512 > $O/c_i8_b1024_d512.json 2>> $O/c.err
tail -n 5 $O/c.err
python tools/summarize.py $O/c_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*" | paste - - - - -
