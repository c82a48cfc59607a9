# usage: gpu_r2_multi.sh N   (under gpurun --gpus N): multi-rank gather equality test, bench weak / strong, gather transports
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/m${N}_pytest.log; tail -5 gpurun_out/m${N}_pytest.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench_weak_p2p.json 2> gpurun_out/m${N}_bench_weak_p2p.err; tail -2 gpurun_out/m${N}_bench_weak_p2p.err
echo "== weak p2p"; python tools/show_bench.py gpurun_out/m${N}_bench_weak_p2p.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --gather nccl --no-components --no-alt-falloff > gpurun_out/m${N}_bench_weak_nccl.json 2> gpurun_out/m${N}_bench_weak_nccl.err; tail -2 gpurun_out/m${N}_bench_weak_nccl.err
echo "== weak nccl"; python tools/show_bench.py gpurun_out/m${N}_bench_weak_nccl.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --scaling strong --no-components --no-alt-falloff > gpurun_out/m${N}_bench_strong_p2p.json 2> gpurun_out/m${N}_bench_strong_p2p.err; tail -2 gpurun_out/m${N}_bench_strong_p2p.err
echo "== strong p2p"; python tools/show_bench.py gpurun_out/m${N}_bench_strong_p2p.json
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-numa --no-components --no-alt-falloff > gpurun_out/m${N}_bench_weak_nonuma.json 2> gpurun_out/m${N}_bench_weak_nonuma.err
echo "== weak p2p, no NUMA binding"; python tools/show_bench.py gpurun_out/m${N}_bench_weak_nonuma.json
timeout 300 $TR tools/pcie_probe.py > gpurun_out/m${N}_pcie.log 2>&1; grep "H2D" gpurun_out/m${N}_pcie.log
nvidia-smi topo -m > gpurun_out/m${N}_topo.log 2>&1; head -14 gpurun_out/m${N}_topo.log
