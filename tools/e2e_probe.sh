for B in 32 128; do
for S in 20 200; do
echo "== batch $B steps $S"
BENCH_TRACE=1 LWOP_TRACE_HOST=1 python bench.py --batch $B --steps $S --warmup 5 --no-cpu --no-extras 2> gpurun_out/e2e_probe_${B}_${S}.err | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=b['config']
print('value',round(b['value']),'e2e',round(b['e2e']['value']),'ratio',round(b['e2e']['value']/b['value'],3),'ms/step',round(b['ms_per_step'],4),'e2e wall ms/step',round(c['e2e_wall_ms_per_step'],4),'lanes',c['steps_in_flight_per_gpu'],c['e2e_steps_in_flight_per_gpu'],'clk',b['clocks']['sm_mhz'], 'sync',round(c['e2e_one_synchronous_call_per_step']['value']))"
grep -E "host ms per step|lwop host trace" gpurun_out/e2e_probe_${B}_${S}.err | tail -3
done; done
