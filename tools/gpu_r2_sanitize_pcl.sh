# compute-sanitizer over the point-cloud / task parity tests only (memcheck + racecheck); logs -> gpurun_out/s_*.log
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_pcl_gpu.py tests/test_task_gpu.py -m gpu -x -q -k "mixed_sizes or golden_reference or cluster or reset_then or multi_region or task_observation" > gpurun_out/s_memcheck_pcl_task.log 2>&1
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_pcl_gpu.py tests/test_task_gpu.py -m gpu -x -q -k "mixed_sizes or cluster or multi_region or reset_then" > gpurun_out/s_racecheck_pcl_task.log 2>&1
for f in gpurun_out/s_memcheck_pcl_task.log gpurun_out/s_racecheck_pcl_task.log; do echo == $f; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|hazard" $f | sort | uniq -c | head -8; done
