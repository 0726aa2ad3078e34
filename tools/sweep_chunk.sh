for chunk in 148 296 444 592 1024 4096; do
  for thr in 320 480 608; do
    PSXB200_PACK_THREADS=$thr python bench.py --steps 30 --no-cpu --chunk $chunk 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('chunk=$chunk thr=$thr value=%.0f ms=%.3f pack_ms/launch=%.4f share=%s e2e=%.0f'%(d['value'],d['ms_per_step'],d['roofline']['launch_ms'],{k:round(v,2) for k,v in d['roofline']['kernel_share'].items()},d['e2e']['value']))"
  done
done
