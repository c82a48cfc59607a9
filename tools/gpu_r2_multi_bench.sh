# usage: gpu_r2_multi_bench.sh N : weak (copy-engine gather) and strong scaling bench lines only
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-alt-falloff > gpurun_out/mb${N}_bench_weak_p2p.json 2> gpurun_out/mb${N}_bench_weak_p2p.err; tail -2 gpurun_out/mb${N}_bench_weak_p2p.err
echo "== weak p2p"; python tools/show_bench.py gpurun_out/mb${N}_bench_weak_p2p.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-components --no-alt-falloff --no-e2e > gpurun_out/mb${N}_bench_strong_p2p.json 2> gpurun_out/mb${N}_bench_strong_p2p.err; tail -2 gpurun_out/mb${N}_bench_strong_p2p.err
echo "== strong p2p"; python tools/show_bench.py gpurun_out/mb${N}_bench_strong_p2p.json
