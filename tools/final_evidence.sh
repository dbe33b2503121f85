set -u
OUT=gpurun_out
python bench.py --steps 20 --warmup 5 > $OUT/r2f_bench_1gpu.json 2> $OUT/r2f_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r2f_bench_reference.json 2> $OUT/r2f_bench_reference.err; echo "ref rc=$?"
python tools/time_draw.py 256 4 > $OUT/r2f_draw.jsonl 2>&1; PROBE_W=656 python tools/time_draw.py 64 20 >> $OUT/r2f_draw.jsonl 2>&1
python tools/rough_weights.py 184 184 2 > $OUT/r2f_rough_weights.jsonl 2>&1; python tools/rough_weights.py 368 368 1 >> $OUT/r2f_rough_weights.jsonl 2>&1
# launch list of a short bench run (serialised, cold-cache: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extras > $OUT/r2f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
# ncu --set full of the rasteriser kernels
python tools/time_draw.py 256 4 > /dev/null 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:draw_mark_kernel' -s 2 -c 1 -o $OUT/r2f_draw_mark -f python tools/time_draw.py 256 4 > $OUT/ncu_draw_mark.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:draw_resolve4_kernel' -s 2 -c 1 -o $OUT/r2f_draw_resolve -f python tools/time_draw.py 256 4 > $OUT/ncu_draw_resolve.log 2>&1
ls -la $OUT/r2f_*
