# fill split variants: which kernel writes which part of the no-contact result (1 colour, 2 gel_depth, 4 obs); value = geom parts + 1
mkdir -p gpurun_out
for v in 6 5 2 1 8; do
  IGI_FILL_SPLIT=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-alt-falloff > gpurun_out/fs_$v.json 2> gpurun_out/fs_$v.err
  echo "== IGI_FILL_SPLIT=$v (geom parts $((v-1)))"; python tools/show_bench.py gpurun_out/fs_$v.json | head -3
done
