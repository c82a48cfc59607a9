python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_ref.json 2>> gpurun_out/r1_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r1_ncu_bench.log 2>&1
tail -3 gpurun_out/r1_pytest.log; cat gpurun_out/r1_bench.json; tail -3 gpurun_out/r1_bench.err
