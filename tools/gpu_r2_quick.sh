# quick perf + parity check of the tactile kernels: tactile / task / full-size parity tests, then bench lines
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tactile_gpu.py tests/test_task_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/q_pytest.log
tail -4 gpurun_out/q_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; tail -2 gpurun_out/q_bench.err
python tools/show_bench.py gpurun_out/q_bench.json
for extra in "$@"; do
  name=$(echo "$extra" | tr -c 'a-zA-Z0-9' '_')
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e $extra > gpurun_out/q_bench_$name.json 2> gpurun_out/q_bench_$name.err
  tail -2 gpurun_out/q_bench_$name.err; echo "== $extra"; python tools/show_bench.py gpurun_out/q_bench_$name.json
done
