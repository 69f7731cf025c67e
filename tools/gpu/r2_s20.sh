set -x
timeout 400 python -m pytest tests/test_gpu_live.py -x -q --tb=short -p no:cacheprovider --timeout 100 > gpurun_out/r2s20_tests.log 2>&1
tail -15 gpurun_out/r2s20_tests.log
B="timeout 120 python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 30"
$B --rows 1000000 > gpurun_out/r2s20_f32_b256_1M.json 2> gpurun_out/r2s20.err
$B --rows 1000000 --opt guess=0 > gpurun_out/r2s20_f32_b256_1M_noguess.json 2>> gpurun_out/r2s20.err
$B --rows 1250000 > gpurun_out/r2s20_f32_b256_shard.json 2>> gpurun_out/r2s20.err
$B --rows 500000 > gpurun_out/r2s20_f32_b256_500k.json 2>> gpurun_out/r2s20.err
$B --rows 1250000 --dtype i8 --batch 1024 > gpurun_out/r2s20_i8_shard.json 2>> gpurun_out/r2s20.err
$B --rows 1000000 --batch 16 > gpurun_out/r2s20_f32_b16_1M.json 2>> gpurun_out/r2s20.err
$B --rows 1000000 --batch 16 --opt guess=0 > gpurun_out/r2s20_f32_b16_1M_noguess.json 2>> gpurun_out/r2s20.err
$B --opt guess_max_rows=20000000 > gpurun_out/r2s20_f32_b256_10M_guess.json 2>> gpurun_out/r2s20.err
tail -3 gpurun_out/r2s20.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s20_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
