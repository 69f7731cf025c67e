set -x
O=gpurun_out
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 5 python __graft_entry__.py smoke > $O/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?"
tail -n 12 $O/sanitizer_memcheck.log
timeout 2400 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 > $O/r01_final_pytest_gpu.log 2>&1
tail -4 $O/r01_final_pytest_gpu.log
timeout 600 python bench.py > $O/r01_final_f32_b256.json 2> $O/f.err
timeout 600 python bench.py --dtype i8 --batch 1024 > $O/r01_final_i8_b1024.json 2>> $O/f.err
B="timeout 300 python bench.py --no-cpu"
$B --dim 512 > $O/r01_final_f32_b256_d512.json 2>> $O/f.err
$B --dtype i8 --batch 1024 --dim 512 > $O/r01_final_i8_b1024_d512.json 2>> $O/f.err
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 10 > $O/r01_final_f16_b4096_shard.json 2>> $O/f.err
tail -n 3 $O/f.err
python tools/summarize.py $O/r01_final_f32_b256.json $O/r01_final_i8_b1024.json $O/r01_final_f32_b256_d512.json $O/r01_final_i8_b1024_d512.json $O/r01_final_f16_b4096_shard.json | grep -o "^[^ ]*\|qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*\|parity {[^}]*}" | paste - - - - - -
