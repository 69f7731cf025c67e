# usage: bash tools/gpu/run_multi.sh N   (under gpurun --gpus N)
set -x
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/multi_f32_b256_n$N.json 2> gpurun_out/multi_err_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/multi_i8_b1024_n$N.json 2>> gpurun_out/multi_err_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/gpu/multi_parity.py > gpurun_out/multi_parity_n$N.log 2>&1
tail -5 gpurun_out/multi_err_n$N.log gpurun_out/multi_parity_n$N.log
cat gpurun_out/multi_f32_b256_n$N.json gpurun_out/multi_i8_b1024_n$N.json
