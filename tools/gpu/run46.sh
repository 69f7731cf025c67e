set -x
O=gpurun_out
timeout 2000 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider --timeout 900 > $O/r01_final_pytest_gpu.log 2>&1
tail -3 $O/r01_final_pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py > $O/r01_final_f32_b256.json 2> $O/f.err
tail -n 2 $O/f.err
python tools/summarize.py $O/r01_final_f32_b256.json | grep -o "qps *[0-9]*\|e2e *[0-9]*\|scan_ms *[0-9.]*\|frac [0-9.]*\|parity {[^}]*}"
grep -o '"full_size_properties": {[^}]*}' $O/r01_final_f32_b256.json
