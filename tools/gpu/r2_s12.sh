set -x
timeout 400 python -m pytest tests/test_gpu_live.py tests/test_gpu_tc_f32.py -x -q --tb=short -p no:cacheprovider --timeout 100 > gpurun_out/r2s12_tests.log 2>&1
tail -5 gpurun_out/r2s12_tests.log
B="timeout 120 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s12_f32_b256_live.json 2> gpurun_out/r2s12.err
$B --opt live=0 > gpurun_out/r2s12_f32_b256_chunkdefer.json 2>> gpurun_out/r2s12.err
$B --opt live=0 --opt chunk_growth_x100=250 > gpurun_out/r2s12_f32_b256_chunkdefer_g25.json 2>> gpurun_out/r2s12.err
$B --opt live=0 --opt img8_defer=0 > gpurun_out/r2s12_f32_b256_r1.json 2>> gpurun_out/r2s12.err
$B --rows 1000000 > gpurun_out/r2s12_f32_b256_1M_live.json 2>> gpurun_out/r2s12.err
$B --rows 1000000 --opt live=0 > gpurun_out/r2s12_f32_b256_1M_chunkdefer.json 2>> gpurun_out/r2s12.err
$B --rows 1000000 --opt live=0 --opt img8_defer=0 > gpurun_out/r2s12_f32_b256_1M_r1.json 2>> gpurun_out/r2s12.err
$B --batch 1024 --opt live=0 > gpurun_out/r2s12_f32_b1024_chunkdefer.json 2>> gpurun_out/r2s12.err
$B --batch 128 --opt live=0 > gpurun_out/r2s12_f32_b128_chunkdefer.json 2>> gpurun_out/r2s12.err
$B --batch 1 --opt live=0 > gpurun_out/r2s12_f32_b1_chunkdefer.json 2>> gpurun_out/r2s12.err
tail -5 gpurun_out/r2s12.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s12_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'refr', round(st.get('live_refreshes_per_step',0)), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
