# usage: bash tools/sweep_cfg.sh "thr:minctas thr:minctas ..." "chunk chunk ..."
for cfg in $1; do
  thr=${cfg%%:*}; mc=${cfg##*:}
  for chunk in $2; do
    PSXB200_PACK_THREADS=$thr PSXB200_PACK_MIN_CTAS=$mc python bench.py --steps 30 --no-cpu --chunk $chunk 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('thr=$thr minctas=$mc chunk=$chunk value=%.0f ms=%.3f pack_ms/launch=%.4f share=%s e2e=%.0f'%(d['value'],d['ms_per_step'],d['roofline']['launch_ms'],{k:round(v,2) for k,v in d['roofline']['kernel_share'].items()},d['e2e']['value']))"
  done
done
