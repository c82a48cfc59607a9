# session-8 call A: verify HEAD on the GPU: parity tests, smoke, bench (+reference arm), ncu launch list
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) 2>&1 | tail -15 > gpurun_out/a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2>> gpurun_out/a_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/a_ncu_bench.log 2>&1
tail -6 gpurun_out/a_pytest.log; tail -2 gpurun_out/a_smoke.log; cat gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
python tools/launches.py gpurun_out/a_launches.csv 12
