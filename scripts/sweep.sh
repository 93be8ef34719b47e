#!/bin/bash
# dev tool: sweep search-kernel launch knobs (env) on the 1M config; prints one summary line each
run() {
  env "$@" timeout 150 python bench.py --steps 10 --latent 24 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$*', 'qps', round(d['value']), 'recall', round(d['recall_at_10'],4), 'frac', round(r['frac'],3), 'kernel_ms', round(r['kernel_ms'],3), 'e2e', round(d['e2e']['value']))
except Exception as e:
    print('$*', 'FAILED', e)
"
}
for args in "$@"; do run $args; done
