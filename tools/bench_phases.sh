#!/bin/bash
# prints the phase times of a short bench run: tools/bench_phases.sh [extra bench args]
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
p=d['phases_ms']
print('%s | %.3f Grays/s  step %.3f ms | build %.3f fwd %.3f loss %.3f bwd %.3f | frac %.3f' % (d['config']['workload'][:14], d['value']/1e9, d['ms_per_step'], p['bvh_build'], p['fwd'], p['loss_grad'], p['bwd'], d['roofline']['frac']))"
