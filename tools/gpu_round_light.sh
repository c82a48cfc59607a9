# light check of HEAD: GPU tests, smoke, bench (+reference arm), ncu launch list (no --set full captures)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/f_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2>> gpurun_out/f_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/f_ncu_bench.log 2>&1
tail -3 gpurun_out/f_pytest.log; tail -2 gpurun_out/f_smoke.log; tail -3 gpurun_out/f_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print("ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["ms_per_step"],3), "launches", d["gpu_launches"], {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
python tools/launches.py gpurun_out/f_launches.csv 7
