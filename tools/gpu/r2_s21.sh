# guess-start sweep (10M / 1M / shard), int8 with a guessed start, and full ncu captures of the live launches
set -x
B="timeout 150 python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 30"
E=gpurun_out/r2s21.err
: > $E
$B > gpurun_out/r2s21_f32_10M_default.json 2>> $E
for f in 64 128 256 512; do
  $B --opt guess_max_rows=20000000 --opt guess_factor=$f > gpurun_out/r2s21_f32_10M_g$f.json 2>> $E
done
for f in 10 16 24 40; do
  $B --rows 1000000 --opt guess_factor=$f > gpurun_out/r2s21_f32_1M_g$f.json 2>> $E
done
for f in 16 32; do
  $B --rows 1250000 --opt guess_factor=$f > gpurun_out/r2s21_f32_shard_g$f.json 2>> $E
done
$B --rows 2500000 > gpurun_out/r2s21_f32_2p5M_default.json 2>> $E
$B --rows 2500000 --opt guess_max_rows=20000000 --opt guess_factor=32 > gpurun_out/r2s21_f32_2p5M_g32.json 2>> $E
$B --rows 5000000 > gpurun_out/r2s21_f32_5M_default.json 2>> $E
$B --rows 5000000 --opt guess_max_rows=20000000 --opt guess_factor=64 > gpurun_out/r2s21_f32_5M_g64.json 2>> $E
$B --dtype i8 --batch 1024 --steps 15 > gpurun_out/r2s21_i8_10M_default.json 2>> $E
$B --dtype i8 --batch 1024 --steps 15 --opt guess_max_rows=20000000 --opt guess_factor=128 > gpurun_out/r2s21_i8_10M_g128.json 2>> $E
$B --batch 1024 --steps 10 > gpurun_out/r2s21_f32_b1024_default.json 2>> $E
$B --batch 1024 --steps 10 --opt guess_max_rows=20000000 --opt guess_factor=128 > gpurun_out/r2s21_f32_b1024_g128.json 2>> $E
tail -5 $E
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s21_*.json')):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1]); r=d['roofline']; st=d.get('search_stats',{})
        print(f.split('/')[-1][7:-5], round(d['value']), round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'L/step', d['gpu_launches']/d['steps'], 'ovf', d.get('overflow_rescans'), 'resc/q', round(st.get('rescored_rows_per_query',0)), 'defer/q', round(st.get('deferred_rows_per_query',0)), d['full_size_properties'].get('sampled_rows_beating_kth'))
    except Exception as e: print(f, 'ERR', e)
PY
NCU="ncu --clock-control none"
timeout 400 $NCU --set full --import-source on -k regex:scan_img8 -s 3 -c 1 -o gpurun_out/r2s21_ncu_f32_10M_live python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 > /dev/null 2>&1
timeout 400 $NCU --set full --import-source on -k regex:scan_img8 -s 3 -c 1 -o gpurun_out/r2s21_ncu_f32_10M_guess python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 --opt guess_max_rows=20000000 --opt guess_factor=128 > /dev/null 2>&1
timeout 400 $NCU --set full --import-source on -k regex:scan_img8 -s 3 -c 1 -o gpurun_out/r2s21_ncu_f32_1M_guess python bench.py --no-cpu --no-configs --sustain-seconds 0 --steps 1 --rows 1000000 > /dev/null 2>&1
ls -la gpurun_out | grep r2s21_ncu
