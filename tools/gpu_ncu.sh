# ncu --set full capture of one launch of each hot kernel (bench workload, 4096 envs)
set -x
for k in tac_contact tac_geom fps_warp_kernel pcl_compact_kernel; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
      -f -o gpurun_out/ncu_$k python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
