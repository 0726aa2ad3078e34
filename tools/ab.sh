# usage: bash tools/ab.sh [-w "<bench args>"] variantA variantB ...  (files variants/<name>.so); two alternating rounds each
extra=""
if [ "$1" = "-w" ]; then extra="$2"; shift 2; fi
for round in 1 2; do for v in "$@"; do PSXB200_LIB=$PWD/variants/$v.so python bench.py --steps 50 --no-cpu $extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('%-10s value=%.0f ms=%.3f dct=%.4f pack=%.4f'%('$v',d['value'],d['ms_per_step'],d['roofline']['kernel_ms_total']*d['roofline']['kernel_share']['bs_dct_kernel']/d['steps'],d['roofline']['launch_ms']))"; done; done
