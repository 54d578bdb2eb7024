mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bvh.py -m gpu -q -x -s > gpurun_out/bvh_pytest.log 2>&1; tail -6 gpurun_out/bvh_pytest.log | cut -c1-300
timeout 300 python tools/bvh_build_time.py > gpurun_out/bvh_time.json 2>&1; cat gpurun_out/bvh_time.json | cut -c1-1500
