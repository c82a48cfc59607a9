# call D: full GPU suite (new student stages, 4-pixel pcl_compact, look-ahead TMA zero fill), bench, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/d_pytest.log; cat gpurun_out/d_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; tail -3 gpurun_out/d_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d_bench.json'))
print("ms/step",round(d["ms_per_step"],3),"e2e",d.get("e2e",{}).get("ms_per_step"), {k:round(v["ms"],3) for k,v in d["kernels"].items()})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python tools/launches.py gpurun_out/d_launches.csv 8
