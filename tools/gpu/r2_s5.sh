set -x
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_img8 -s 2 -c 1 -o gpurun_out/r2s5_live_img8 python bench.py --no-cpu --steps 1 > gpurun_out/r2s5_ncu.log 2>&1
tail -3 gpurun_out/r2s5_ncu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:pkv -c 60 --csv --log-file gpurun_out/r2s5_launches.csv python bench.py --no-cpu --steps 1 > gpurun_out/r2s5_ncu2.log 2>&1
tail -3 gpurun_out/r2s5_ncu2.log
ls -la gpurun_out/ | tail -5
