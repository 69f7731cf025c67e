set -x
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 0 -c 60 --csv --log-file gpurun_out/r2s8_launches_1M.csv python bench.py --no-cpu --steps 1 --rows 1250000 > gpurun_out/r2s8_ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"scan_|rescore|select|prep_|reset_|finalize" -s 0 -c 45 --csv --log-file gpurun_out/r2s8_launches_10M.csv python bench.py --no-cpu --steps 1 > gpurun_out/r2s8_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_img8 -s 2 -c 1 -o gpurun_out/r2s8_live_img8 python bench.py --no-cpu --steps 1 --rows 1250000 > gpurun_out/r2s8_ncu3.log 2>&1
ls -la gpurun_out | tail -4
