# round-2 check of the working tree: GPU tests, smoke, bench (default config), optional extra bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/c_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c_smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 $BENCH_EXTRA > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err
tail -15 gpurun_out/c_pytest.log; tail -2 gpurun_out/c_smoke.log; tail -3 gpurun_out/c_bench.err
python tools/show_bench.py gpurun_out/c_bench.json
for extra in "$@"; do
  name=$(echo "$extra" | tr -c 'a-zA-Z0-9' '_')
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $extra > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  tail -2 gpurun_out/c_bench_$name.err; echo "== $extra"; python tools/show_bench.py gpurun_out/c_bench_$name.json
done
