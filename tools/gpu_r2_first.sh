# round-2 first call: sanity of HEAD, copy-engine fill probe, prefill experiment, env sweep, sanitizer pass
mkdir -p gpurun_out
IGI_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/a_pytest.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ce_fill_probe tools/ce_fill_probe.cu && timeout 120 /tmp/ce_fill_probe > gpurun_out/a_ce_probe.log 2>&1
for e in 4096 256 1024 16384; do
  timeout 600 python bench.py --envs $e --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/a_bench_$e.json 2> gpurun_out/a_bench_$e.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --prefill-gel-depth > gpurun_out/a_bench_prefill.json 2> gpurun_out/a_bench_prefill.err
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_tactile_gpu.py -m gpu -x -q -k "batched_render or update_mask or coloured" > gpurun_out/a_memcheck_tactile.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_tactile_gpu.py -m gpu -x -q -k "batched_render" > gpurun_out/a_racecheck_tactile.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_pcl_gpu.py -m gpu -x -q -k "mixed_sizes or golden_reference or cluster" > gpurun_out/a_memcheck_pcl.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_pcl_gpu.py -m gpu -x -q -k "mixed_sizes or cluster" > gpurun_out/a_racecheck_pcl.log 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_ce_probe.log
python - <<'PY'
import json
for e in (4096, 256, 1024, 16384, "prefill"):
    try:
        d=json.load(open(f'gpurun_out/a_bench_{e}.json'))
        print(e, "ms/step",round(d["ms_per_step"],3),"e2e",round(d.get("e2e",{}).get("ms_per_step",0),3), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
    except Exception as ex:
        print(e, "failed", ex)
PY
for f in gpurun_out/a_memcheck_tactile.log gpurun_out/a_racecheck_tactile.log gpurun_out/a_memcheck_pcl.log gpurun_out/a_racecheck_pcl.log; do echo == $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" $f | sort | uniq -c | head -8; done
