set -x
timeout 900 python -m pytest tests/test_gpu_live.py tests/test_gpu_tc.py -q --tb=short -p no:cacheprovider --timeout 300 > gpurun_out/r2s7_live.log 2>&1
tail -12 gpurun_out/r2s7_live.log
B="timeout 300 python bench.py --no-cpu --steps 20"
$B > gpurun_out/r2s7_f32_b256.json 2> gpurun_out/r2s7.err
$B --opt live_start_rows=32768 > gpurun_out/r2s7_f32_b256_ls32k.json 2>> gpurun_out/r2s7.err
$B --opt live_start_rows=262144 > gpurun_out/r2s7_f32_b256_ls256k.json 2>> gpurun_out/r2s7.err
$B --rows 1000000 > gpurun_out/r2s7_f32_b256_1M.json 2>> gpurun_out/r2s7.err
$B --rows 1000000 --opt live_start_rows=32768 > gpurun_out/r2s7_f32_b256_1M_ls32k.json 2>> gpurun_out/r2s7.err
$B --batch 1 > gpurun_out/r2s7_f32_b1.json 2>> gpurun_out/r2s7.err
$B --batch 128 > gpurun_out/r2s7_f32_b128.json 2>> gpurun_out/r2s7.err
$B --batch 1024 > gpurun_out/r2s7_f32_b1024.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 1024 --steps 10 > gpurun_out/r2s7_i8_b1024.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 1024 --steps 10 --opt live=0 > gpurun_out/r2s7_i8_b1024_chunked.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 128 > gpurun_out/r2s7_i8_b128.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 128 --opt live=0 > gpurun_out/r2s7_i8_b128_chunked.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 1024 --rows 1250000 > gpurun_out/r2s7_i8_b1024_shard.json 2>> gpurun_out/r2s7.err
$B --dtype i8 --batch 1024 --rows 1250000 --opt live=0 > gpurun_out/r2s7_i8_b1024_shard_chunked.json 2>> gpurun_out/r2s7.err
$B --rows 1250000 > gpurun_out/r2s7_f32_b256_shard.json 2>> gpurun_out/r2s7.err
$B --rows 1250000 --opt live=0 > gpurun_out/r2s7_f32_b256_shard_chunked.json 2>> gpurun_out/r2s7.err
tail -5 gpurun_out/r2s7.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s7_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][5:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'refr', round(st.get('live_refreshes_per_step',0)), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
