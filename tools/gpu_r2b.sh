# call B: TMA zero fill in tac_contact (default split 5) against the all-in-geom split 7; ncu launch list
mkdir -p gpurun_out
bash tools/gpu_variants2.sh "" "-DFILL_GEOM_PARTS=7" "-DFILL_GEOM_PARTS=1"
python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python tools/launches.py gpurun_out/b_launches.csv 12
