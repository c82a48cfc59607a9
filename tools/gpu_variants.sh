# builds tactile variants ON the box (IGI_NVCC_EXTRA) and prints the bench summary of each; the last build is the default one
mkdir -p gpurun_out
for v in "$@"; do
  IGI_NVCC_EXTRA="$v" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-alt-falloff > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; tail -2 gpurun_out/v_bench.err
  echo "== variant '$v'"; python tools/show_bench.py gpurun_out/v_bench.json 2>/dev/null | sed -n 1,3p
done
python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1
