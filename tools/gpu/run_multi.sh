# usage: bash tools/gpu/run_multi.sh N   (under gpurun --gpus N)
set -x
N=${1:-2}
nvidia-smi -L
TR="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29513 tools/gpu/multi_parity.py > gpurun_out/r2_multi_parity_n$N.log 2>&1
tail -6 gpurun_out/r2_multi_parity_n$N.log
$TR --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_multi_f32_b256_n$N.json 2> gpurun_out/r2_multi_err_n$N.log
$TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/r2_multi_i8_b1024_n$N.json 2>> gpurun_out/r2_multi_err_n$N.log
if [ "$N" = "8" ]; then
  # BASELINE config 4: 50M x 512 f16, batch 4096, row-sharded over 8 GPUs; config 5: vector top-k AND tag bitmap, 10M, 8 GPUs
  $TR --master-port 29514 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --dtype f16 --dim 512 --rows 50000000 --batch 4096 --sustain-seconds 0 > gpurun_out/r2_multi_c4_f16_b4096_n8.json 2>> gpurun_out/r2_multi_err_n$N.log
  $TR --master-port 29515 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu --bitmap-density 0.1 --sustain-seconds 0 > gpurun_out/r2_multi_c5_f32_bitmap10_n8.json 2>> gpurun_out/r2_multi_err_n$N.log
  $TR --master-port 29516 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu --dtype i8 --batch 1024 --bitmap-density 0.5 --sustain-seconds 0 > gpurun_out/r2_multi_c5_i8_bitmap50_n8.json 2>> gpurun_out/r2_multi_err_n$N.log
fi
tail -5 gpurun_out/r2_multi_err_n$N.log
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_multi_*_n$N.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f.split('/')[-1], round(d['value']), 'ms', round(d['ms_per_step'],3), 'kern', round(r['kernel_ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'sust', d['sustained'] and round(d['sustained']['value']), 'L', d['gpu_launches']/d['steps'])
    except Exception as e: print(f,'ERR',e)
PY
