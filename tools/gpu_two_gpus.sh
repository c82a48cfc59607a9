# 2 GPUs: NCCL gather bit-equality, bench under torchrun (both arms), world-2 value
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR tools/check_gather_nccl.py 64 > gpurun_out/n2_gather.log 2>&1; grep -v Warn gpurun_out/n2_gather.log | grep -i "nccl gather\|Error\|assert" | head -5
$TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; tail -2 gpurun_out/n2_bench.err; head -c 1500 gpurun_out/n2_bench.json; echo
$TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2_ref.json 2>> gpurun_out/n2_bench.err; head -c 400 gpurun_out/n2_ref.json; echo
