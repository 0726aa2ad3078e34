# last gpurun call of a round: the driver's own sequence (tests, smoke, bench both arms) + the FDCT capture
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/v8_tests.log
python __graft_entry__.py smoke > gpurun_out/v8_smoke.log 2>&1
python bench.py > gpurun_out/v8_bench.json 2> gpurun_out/v8_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/v8_bench_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/v8_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:bs_dct_kernel -s 3 -c 1 -o gpurun_out/v8_dct python bench.py --steps 3 --warmup 2 --no-cpu > /dev/null 2>&1
cat gpurun_out/v8_tests.log gpurun_out/v8_smoke.log
