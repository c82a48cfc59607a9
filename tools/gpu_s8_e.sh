mkdir -p gpurun_out
python -m pytest tests/test_traj_gpu.py tests/test_task_gpu.py -m gpu -q 2>&1 | tail -12
python tools/pcie_probe.py 2>&1 | tee gpurun_out/e_pcie.txt
