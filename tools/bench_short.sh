#!/bin/bash
# usage: tools/bench_short.sh <extra bench args...>  -> prints value, roofline frac, window
timeout -s KILL 120 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu "$@" 2>&1 | tail -1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); print('ARGS', ' '.join(sys.argv[1:]), '| samples/s %.0f' % d['value'], '| frac %.4f' % d['roofline']['frac'], '| ms %.3f' % d['roofline']['kernel_ms'], '| win', d['config']['bev_window_cells'], '| thr', d['config']['threads'])
except Exception as e: print('FAILED', e)
" "$@"
