#!/bin/bash
# e2e (host-buffer path) against value (device-resident) at small slices on ONE GPU, 20 and 200 steps, with the pool's
# uploads through one shared copy stream (default) and through a private stream per engine (LWOP_POOL_SHARED_COPY=0)
for SH in 1 0; do for B in 32 128; do for S in 20 200; do
LWOP_POOL_SHARED_COPY=$SH python bench.py --batch $B --steps $S --warmup 5 --no-cpu --no-extras 2> /dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=b['config']
print('shared_copy $SH batch $B steps $S value',round(b['value']),'e2e',round(b['e2e']['value']),'ratio',round(b['e2e']['value']/b['value'],3),'ms/step',round(b['ms_per_step'],4),'e2e wall ms/step',round(c['e2e_wall_ms_per_step'],4),'clk',b['clocks']['sm_mhz'])"
done; done; done
