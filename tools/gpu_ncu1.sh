# usage: gpu_ncu1.sh <kernel regex> <tag>  : one ncu --set full capture from the bench workload
k=$1; tag=$2
ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 2 --launch-count 1 \
    -f -o gpurun_out/ncu_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
