# A/B differently built libraries (ab/*.so) on the default bench workload: kernel ms per batch
for f in ab/*.so; do
  CTTS_GPU_LIB=$PWD/$f timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e ${AB_ARGS} 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$f', round(d['ms_per_step'],3))"
done
