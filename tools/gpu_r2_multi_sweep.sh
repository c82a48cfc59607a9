# usage: gpu_r2_multi_sweep.sh N : BASELINE config 5 (envs-per-GPU sweep) on N GPUs, weak scaling, copy-engine gather
N=$1
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --config sweep --no-cpu-baseline --no-e2e --no-alt-falloff --no-components > gpurun_out/ms${N}_bench_sweep.json 2> gpurun_out/ms${N}_bench_sweep.err; tail -3 gpurun_out/ms${N}_bench_sweep.err
python tools/show_bench.py gpurun_out/ms${N}_bench_sweep.json
