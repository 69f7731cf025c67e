# s23 epilogue (no spills) + L2 prefetch + batched gathers + guessed start from 64k rows: tests, benches, one capture
set -x
timeout 900 python -m pytest tests/test_gpu_live.py tests/test_gpu_at_size.py -x -q --tb=short -p no:cacheprovider --timeout 200 > gpurun_out/r2s25_tests.log 2>&1
tail -5 gpurun_out/r2s25_tests.log
B="timeout 150 python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 30"
E=gpurun_out/r2s25.err
: > $E
$B > gpurun_out/r2s25_f32_10M.json 2>> $E
$B --rows 1000000 > gpurun_out/r2s25_f32_1M.json 2>> $E
$B --rows 1250000 > gpurun_out/r2s25_f32_shard.json 2>> $E
$B --rows 500000 > gpurun_out/r2s25_f32_500k.json 2>> $E
$B --rows 100000 > gpurun_out/r2s25_f32_100k.json 2>> $E
$B --opt guess=0 > gpurun_out/r2s25_f32_10M_noguess.json 2>> $E
$B --batch 1024 --steps 10 > gpurun_out/r2s25_f32_b1024.json 2>> $E
$B --dtype i8 --batch 1024 --steps 15 > gpurun_out/r2s25_i8_10M.json 2>> $E
$B --dtype i8 --batch 256 > gpurun_out/r2s25_i8_b256.json 2>> $E
$B --dtype f16 --dim 512 --rows 6250000 --batch 4096 --steps 5 > gpurun_out/r2s25_f16_shard.json 2>> $E
tail -5 $E
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s25_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5].ljust(22), round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
NCU="ncu --clock-control none"
timeout 400 $NCU --set full --import-source on -k regex:scan_img8 -s 1 -c 1 -o gpurun_out/r2s25_ncu_f32_10M python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 > /dev/null 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --cache-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 40 -c 24 --csv --log-file gpurun_out/r2s25_launches_1M.csv python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 4 --rows 1000000 > /dev/null 2>&1
ls -la gpurun_out | grep r2s25_ncu
