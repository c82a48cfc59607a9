# session-8 full check of HEAD: GPU parity tests, smoke, bench (+reference arm), ncu launch list, ncu --set full of the hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2>> gpurun_out/f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_bench.log 2>&1
for k in tac_contact tac_geom fps_sorted pcl_compact_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
      -f -o gpurun_out/ncu_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
tail -3 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; cat gpurun_out/f_bench.json; tail -3 gpurun_out/f_bench.err
python tools/launches.py gpurun_out/f_launches.csv 8
