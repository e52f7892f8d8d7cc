cd /root/repo
GA_TEST_PROGRESS=/root/repo/gpurun_out/nccl_progress timeout 480 python -m pytest tests/test_ddp_nccl_gpu.py -x -q -m gpu > gpurun_out/nccl_test.log 2>&1
echo rc=$?
tail -5 gpurun_out/nccl_test.log
cat gpurun_out/nccl_progress.rank0
