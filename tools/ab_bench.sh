#!/bin/bash
# A/B of an environment setting on the bench's device-resident numbers: usage ab_bench.sh "<ENV=.. ENV=..>" [batches...]
# prints value / one-step-at-a-time / e2e for the default environment and for the given one, alternating twice
SETTING=$1; shift
for B in "${@:-256}"; do for rep in 1 2; do for S in "" "$SETTING"; do
env $S python bench.py --batch $B --steps 40 --warmup 5 --no-cpu --no-extras 2>/dev/null | python -c "
import json,sys
b=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=b['config']
print('batch $B [$S] value',round(b['value']),'serial',round(c['one_step_at_a_time']['value']),'e2e',round(b['e2e']['value']),'ms/step',round(b['ms_per_step'],4),'clk',b['clocks']['sm_mhz'])"
done; done; done
