set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/v8_tests.log
python bench.py > gpurun_out/v8_bench.json 2> gpurun_out/v8_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v8_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v8_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bs_pack_kernel -s 3 -c 1 -o gpurun_out/v8_pack python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bs_dct_kernel -s 3 -c 1 -o gpurun_out/v8_dct python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:adpcm_spu_kernel -s 3 -c 1 -o gpurun_out/v8_spu python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
(bash tools/bench_summary.sh; python tools/bench_adpcm.py; python tools/bench_dropin.py) > gpurun_out/v8_workloads.txt 2>&1
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_bs.py tests/test_gpu_adpcm.py -x -q -k "not full_size and not large_budget and not sbs_config and not 640x512" > gpurun_out/v8_memcheck.log 2>&1
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_bs.py -x -q -k "kat and 64" > gpurun_out/v8_racecheck.log 2>&1
tail -3 gpurun_out/v8_tests.log gpurun_out/v8_memcheck.log gpurun_out/v8_racecheck.log
