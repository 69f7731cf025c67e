set -x
O=gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_f32.py -q --tb=short -p no:cacheprovider --timeout 600 -x > $O/pytest_r35.log 2>&1
tail -5 $O/pytest_r35.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > $O/b_f32_b256.json 2> $O/b.err
$B --batch 1024 > $O/b_f32_b1024.json 2>> $O/b.err
$B --batch 128 > $O/b_f32_b128.json 2>> $O/b.err
$B --batch 16 > $O/b_f32_b16.json 2>> $O/b.err
$B --dtype i8 --batch 1024 > $O/b_i8_b1024.json 2>> $O/b.err
$B --dtype i8 --batch 256 > $O/b_i8_b256.json 2>> $O/b.err
for g in 75 100; do
$B --opt chunk_growth_x100=$g > $O/b_f32_b256_g$g.json 2>> $O/b.err
$B --batch 1024 --opt chunk_growth_x100=$g > $O/b_f32_b1024_g$g.json 2>> $O/b.err
done
tail -n 5 $O/b.err
python tools/summarize.py $O/b_*.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*" | paste - - - - -
