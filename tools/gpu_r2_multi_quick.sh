# usage: gpu_r2_multi_quick.sh N : gather equality test + weak p2p bench line
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/mq${N}_pytest.log; tail -3 gpurun_out/mq${N}_pytest.log
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-components --no-alt-falloff > gpurun_out/mq${N}_bench_weak_p2p.json 2> gpurun_out/mq${N}_bench_weak_p2p.err; tail -2 gpurun_out/mq${N}_bench_weak_p2p.err
python tools/show_bench.py gpurun_out/mq${N}_bench_weak_p2p.json
