set -x
O=gpurun_out
timeout 300 python bench.py --no-cpu --steps 10 > $O/g_f32.json 2> $O/g.err
timeout 300 python bench.py --no-cpu --steps 10 --dtype i8 --batch 1024 > $O/g_i8.json 2>> $O/g.err
timeout 300 python bench.py --no-cpu --steps 10 --metric l2 --batch 16 > $O/g_l2.json 2>> $O/g.err
timeout 300 python bench.py --no-cpu --steps 10 --bitmap-density 0.1 > $O/g_bm.json 2>> $O/g.err
timeout 300 python bench.py --no-cpu --steps 5 --dtype f16 --dim 512 --rows 6250000 --batch 4096 > $O/g_f16.json 2>> $O/g.err
tail -n 3 $O/g.err
for f in $O/g_*.json; do grep -o '"full_size_properties": {[^}]*}' $f; done
