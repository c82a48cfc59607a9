# experiment: stream priorities of the two halves of compute_observations
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-alt-falloff --no-components $EXTRA > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; tail -2 gpurun_out/o_bench.err
  python tools/show_bench.py gpurun_out/o_bench.json 2>/dev/null | sed -n 1,1p
}
run IGI_SIDE_PRIO=0
run IGI_SIDE_PRIO=0 IGI_TAC_PRIO=-1
run IGI_SIDE_PRIO=0 IGI_TAC_PRIO=0
run IGI_SIDE_PRIO=0 IGI_TAC_PRIO=-1 IGI_FPS_CTAS_PER_SM=4
EXTRA="--config sweep" run IGI_SIDE_PRIO=0
python tools/show_bench.py gpurun_out/o_bench.json
EXTRA="--config sweep" run IGI_SIDE_PRIO=-1
python tools/show_bench.py gpurun_out/o_bench.json
