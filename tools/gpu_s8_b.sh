# session-8 call B: student tests after the mask-lifetime fix; per-phase cycle shares of tac_contact
mkdir -p gpurun_out
python -m pytest tests/test_student_gpu.py tests/test_task_gpu.py -m gpu -q 2>&1 | tail -8
IGI_NVCC_EXTRA="-DCT_PROFILE" python -m isaacgyminsertion_b200.build --force > /dev/null 2>&1 || echo build failed
python tools/ct_prof.py 2>&1 | tee gpurun_out/b_ctprof.txt
