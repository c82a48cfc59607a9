# end-of-round record on one GPU: full GPU test suite, smoke, both bench arms, every BASELINE config, ncu launch list + full captures
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/f_${tag}_pytest.log; tail -3 gpurun_out/f_${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_${tag}_smoke.log 2>&1; tail -2 gpurun_out/f_${tag}_smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_${tag}_bench_reference.json 2> gpurun_out/f_${tag}_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_${tag}_bench.json 2> gpurun_out/f_${tag}_bench.err; tail -2 gpurun_out/f_${tag}_bench.err
python tools/show_bench.py gpurun_out/f_${tag}_bench.json
python -c "import json; d=json.load(open('gpurun_out/f_${tag}_bench_reference.json')); print('reference arm', round(d['value'],1), d['cpu_baseline']['cores'], 'cores')"
for cfg in tactile1024 pcl1024 sweep small; do
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config $cfg > gpurun_out/f_${tag}_bench_$cfg.json 2> gpurun_out/f_${tag}_bench_$cfg.err
  echo "== $cfg"; python tools/show_bench.py gpurun_out/f_${tag}_bench_$cfg.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-falloff > gpurun_out/f_${tag}_ncu_bench.log 2>&1
python tools/launches.py gpurun_out/f_${tag}_launches.csv 12
for k in tac_contact tac_geom pcl_compact_kernel fps_sorted_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
      -f -o gpurun_out/f_${tag}_ncu_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-components --no-alt-falloff > gpurun_out/f_${tag}_ncu_$k.log 2>&1
done
python tools/ncu_summary.py gpurun_out/f_${tag}_ncu_tac_contact.ncu-rep gpurun_out/f_${tag}_ncu_tac_geom.ncu-rep gpurun_out/f_${tag}_ncu_pcl_compact_kernel.ncu-rep gpurun_out/f_${tag}_ncu_fps_sorted_kernel.ncu-rep > gpurun_out/f_${tag}_ncu_full.txt 2>&1
grep -E "dram__bytes|gpu__time_duration|smsp__inst_executed.sum|issue_active" gpurun_out/f_${tag}_ncu_full.txt
