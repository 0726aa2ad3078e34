# usage: bash tools/bench_summary.sh [extra bench args...]  -> one summary line per workload
for w in "--noise 3" "--noise 0" "--noise 6" "--workload sbs"; do python bench.py --steps 30 --no-cpu $w "$@" 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print(d['config']['workload'][:34], d['config']['workload'][-14:], 'q=%.1f value=%.0f ms=%.3f dct=%.3f pack=%.3f e2e=%.0f'%(d['config']['quant_scale_mean'],d['value'],d['ms_per_step'],d['roofline']['kernel_ms_total']*d['roofline']['kernel_share']['bs_dct_kernel']/d['steps'],d['roofline']['kernel_ms_total']*d['roofline']['kernel_share']['bs_pack_kernel']/d['steps'],d['e2e']['value']))"; done
