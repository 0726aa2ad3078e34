# usage: bash tools/ab.sh [-w "<bench args>"] variantA variantB ...  (files variants/<name>.so); two alternating rounds each
extra=""
if [ "$1" = "-w" ]; then extra="$2"; shift 2; fi
for round in 1 2; do for v in "$@"; do PSXB200_LIB=$PWD/variants/$v.so python bench.py --steps 50 --no-cpu --headline-only $extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
k=d['roofline']['per_kernel']
print('%-10s value=%.0f ms=%.4f dct=%.4f pack=%.4f'%('$v',d['value'],d['ms_per_step'],k['bs_dct_kernel']['launch_ms'],k['bs_pack_kernel']['launch_ms']))"; done; done
