set -x
timeout 240 python -m pytest tests/test_gpu_live.py -x -q --tb=short -p no:cacheprovider --timeout 100 > gpurun_out/r2s10_live.log 2>&1
tail -5 gpurun_out/r2s10_live.log
B="timeout 120 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s10_f32_b256.json 2> gpurun_out/r2s10.err
$B --opt live_start_rows=32768 > gpurun_out/r2s10_f32_b256_ls32k.json 2>> gpurun_out/r2s10.err
$B --opt live_start_rows=131072 > gpurun_out/r2s10_f32_b256_ls128k.json 2>> gpurun_out/r2s10.err
$B --rows 1000000 > gpurun_out/r2s10_f32_b256_1M.json 2>> gpurun_out/r2s10.err
$B --rows 1000000 --opt live_start_rows=32768 > gpurun_out/r2s10_f32_b256_1M_ls32k.json 2>> gpurun_out/r2s10.err
$B --rows 1000000 --opt live=0 > gpurun_out/r2s10_f32_b256_1M_chunked.json 2>> gpurun_out/r2s10.err
$B --batch 1 > gpurun_out/r2s10_f32_b1.json 2>> gpurun_out/r2s10.err
$B --batch 128 > gpurun_out/r2s10_f32_b128.json 2>> gpurun_out/r2s10.err
$B --batch 1024 > gpurun_out/r2s10_f32_b1024.json 2>> gpurun_out/r2s10.err
$B --rows 1250000 > gpurun_out/r2s10_f32_b256_shard.json 2>> gpurun_out/r2s10.err
tail -5 gpurun_out/r2s10.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s10_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][5:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'refr', round(st.get('live_refreshes_per_step',0)), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
