# One gpurun call: tests, bench (both arms), launch list, ncu full captures of the hot kernels,
# drop-in latencies. Outputs under gpurun_out/ with the prefix given as $1 (default r2).
P=${1:-r2b}
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${P}_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bs_pack_kernel -s 4 -c 1 -o gpurun_out/${P}_pack python bench.py --steps 3 --warmup 2 --no-cpu --headline-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bs_dct_kernel -s 3 -c 1 -o gpurun_out/${P}_dct python bench.py --steps 3 --warmup 2 --no-cpu --headline-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:adpcm_spu_kernel -s 3 -c 1 -o gpurun_out/${P}_spu python bench.py --steps 3 --warmup 2 --no-cpu --only vagi_x1024 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:adpcm_xa_kernel -s 3 -c 1 -o gpurun_out/${P}_xa python bench.py --steps 3 --warmup 2 --no-cpu --only strcd > /dev/null 2>&1
(python tools/bench_dropin.py; python tools/bench_adpcm.py) > gpurun_out/${P}_workloads.txt 2>&1
tail -3 gpurun_out/${P}_tests.log
cat gpurun_out/${P}_workloads.txt
