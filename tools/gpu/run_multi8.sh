set -x
N=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > gpurun_out/multi_f32_b256_n$N.json 2> gpurun_out/multi_err_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --dtype i8 --batch 1024 > gpurun_out/multi_i8_b1024_n$N.json 2>> gpurun_out/multi_err_n$N.log
tail -n 3 gpurun_out/multi_err_n$N.log
cat gpurun_out/multi_f32_b256_n$N.json gpurun_out/multi_i8_b1024_n$N.json | cut -c1-300
