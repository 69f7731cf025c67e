set -x
B="timeout 120 python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 30"
$B > gpurun_out/r2s19_f32_b256.json 2> gpurun_out/r2s19.err
$B --rows 1000000 > gpurun_out/r2s19_f32_b256_1M.json 2>> gpurun_out/r2s19.err
$B --rows 1000000 --opt live=0 > gpurun_out/r2s19_f32_b256_1M_chunked.json 2>> gpurun_out/r2s19.err
$B --rows 1250000 > gpurun_out/r2s19_f32_b256_shard.json 2>> gpurun_out/r2s19.err
$B --rows 1250000 --opt live=0 > gpurun_out/r2s19_f32_b256_shard_chunked.json 2>> gpurun_out/r2s19.err
$B --rows 500000 --opt live=2 > gpurun_out/r2s19_f32_b256_500k_live.json 2>> gpurun_out/r2s19.err
$B --rows 500000 --opt live=0 > gpurun_out/r2s19_f32_b256_500k_chunked.json 2>> gpurun_out/r2s19.err
$B --rows 1250000 --dtype i8 --batch 1024 > gpurun_out/r2s19_i8_shard.json 2>> gpurun_out/r2s19.err
$B --rows 1250000 --dtype i8 --batch 1024 --opt live=0 > gpurun_out/r2s19_i8_shard_chunked.json 2>> gpurun_out/r2s19.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s19_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)))
    except Exception as e: print(f, 'ERR', e)
PY
