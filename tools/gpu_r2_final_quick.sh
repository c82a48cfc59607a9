# re-record of the bench lines only (after a host-side change): GPU tests of the task, both arms, sweep, small, launch list
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_task_gpu.py tests/test_fullsize_gpu.py -m gpu -q 2>&1 | tail -3
timeout 900 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/f_${tag}_bench_reference.json 2> gpurun_out/f_${tag}_bench_reference.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_${tag}_bench.json 2> gpurun_out/f_${tag}_bench.err; tail -2 gpurun_out/f_${tag}_bench.err
python tools/show_bench.py gpurun_out/f_${tag}_bench.json
python -c "import json; d=json.load(open('gpurun_out/f_${tag}_bench_reference.json')); print('reference arm', round(d['value'],1), d['cpu_baseline'])"
for cfg in sweep small; do
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --config $cfg > gpurun_out/f_${tag}_bench_$cfg.json 2> gpurun_out/f_${tag}_bench_$cfg.err
  echo "== $cfg"; python tools/show_bench.py gpurun_out/f_${tag}_bench_$cfg.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-falloff > gpurun_out/f_${tag}_ncu_bench.log 2>&1
python tools/launches.py gpurun_out/f_${tag}_launches.csv 8
