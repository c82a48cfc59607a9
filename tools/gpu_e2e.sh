# e2e leg with different ring depths / without the pcl side stream
for v in 3 4 6; do
  IGI_PIPE_SLOTS=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e2e_$v.json 2> gpurun_out/e2e_$v.err
  python - $v <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/e2e_{sys.argv[1]}.json'))
print("slots", sys.argv[1], "ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
done
IGI_PIPE_SLOTS=4 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-overlap > gpurun_out/e2e_no.json 2> gpurun_out/e2e_no.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/e2e_no.json'))
print("slots 4 no side stream: ms/step", round(d["ms_per_step"],3), "e2e ms", round(d["e2e"]["ms_per_step"],3))
PY
