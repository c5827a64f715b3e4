for c in 3 4; do
  CTTS_GPU_CTAS_PER_SM=$c timeout 300 python bench.py --utts 2048 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas',$c,'ms',round(d['ms_per_step'],3),'win',d['config']['window_samples'],'smem',d['config']['smem_bytes'])"
done
