for v in "$@"; do echo "== $v"; PSXB200_LIB=$PWD/variants/$v.so python tools/bench_adpcm.py 2>&1 | grep -E "files= *(1|128|1024|4096) |streams= *(1|4096|16384) "; done
