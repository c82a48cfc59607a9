# usage: gpu_r2_profile.sh <tag>: phase shares of tac_contact (both light models), ncu launch list, ncu --set full of the hot kernels
tag=$1
mkdir -p gpurun_out
IGI_NVCC_EXTRA="-DCT_PROFILE" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python tools/ct_prof.py > gpurun_out/p_${tag}_phases_default.txt 2>&1; tail -14 gpurun_out/p_${tag}_phases_default.txt
IGI_FALLOFF=none python tools/ct_prof.py > gpurun_out/p_${tag}_phases_none.txt 2>&1; tail -14 gpurun_out/p_${tag}_phases_none.txt
python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p_${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt-falloff > gpurun_out/p_${tag}_ncu_bench.log 2>&1
python tools/launches.py gpurun_out/p_${tag}_launches.csv 7
for k in tac_contact tac_geom; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
      -f -o gpurun_out/p_${tag}_ncu_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-components --no-alt-falloff > gpurun_out/p_${tag}_ncu_$k.log 2>&1
  tail -1 gpurun_out/p_${tag}_ncu_$k.log
done
ncu --set full --clock-control none --import-source on --kernel-name regex:tac_contact --launch-skip 2 --launch-count 1 \
    -f -o gpurun_out/p_${tag}_ncu_tac_contact_none python bench.py --falloff none --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-components --no-alt-falloff > gpurun_out/p_${tag}_ncu_contact_none.log 2>&1
python tools/ncu_summary.py gpurun_out/p_${tag}_ncu_tac_contact.ncu-rep gpurun_out/p_${tag}_ncu_tac_geom.ncu-rep gpurun_out/p_${tag}_ncu_tac_contact_none.ncu-rep > gpurun_out/p_${tag}_ncu_full.txt 2>&1
cat gpurun_out/p_${tag}_ncu_full.txt | head -90
