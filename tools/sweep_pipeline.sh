# Device-resident strv step: frames per launch x stream pipelining (PSXB200_DEVICE_PIPELINE), and
# DRAM traffic per step for the best ones. Writes gpurun_out/${P}_pipeline.txt
P=${1:-r2}
out=gpurun_out/${P}_pipeline.txt
: > $out
for cfg in "4096 1" "2048 1" "1024 1" "512 1" "256 1" "1024 0" "512 0"; do
  set -- $cfg
  PSXB200_DEVICE_PIPELINE=$2 python bench.py --steps 40 --warmup 5 --no-cpu --headline-only --chunk $1 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $1 pipeline $2: %.0f frames/s, %.4f ms/step, sustained %.0f' % (d['value'], d['ms_per_step'], d['sustained']['value']))" >> $out
done
for cfg in "4096 1" "512 1" "1024 1"; do
  set -- $cfg
  PSXB200_DEVICE_PIPELINE=$2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:bs_ -s 40 -c 16 --csv --log-file gpurun_out/${P}_traffic_$1.csv python bench.py --steps 4 --warmup 3 --no-cpu --headline-only --chunk $1 > /dev/null 2>&1
  python - >> $out <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${P}_traffic_$1.csv")) if len(r)>10 and r[0].isdigit()]
tot={}
for r in rows:
    name=r[4].split('(')[0]; metric=r[-3]; unit=r[-2]; val=float(r[-1].replace(',',''))
    mult={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}.get(unit,1)
    if metric.startswith('dram'): tot[name]=tot.get(name,0)+val*mult
n=len(rows)//3
print("chunk $1: dram bytes over %d captured launches (bs_*): %s" % (n, {k:round(v/1e6,1) for k,v in tot.items()}))
PY
done
cat $out
