#!/bin/bash
# run on the GPU box: end-to-end (host buffers) throughput of bench.py for several pipeline piece sizes
for p in 65536 131072 262144 524288; do
  MJB_HOST_PIECE=$p python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('piece $p value %.4g e2e %.4g ms %.3f' % (r['value'], r['e2e']['value'], r['ms_per_step']))"
done
