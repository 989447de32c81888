cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_pa_noise_edge.py tests/test_intra_edges.py tests/test_abi.py -q -m gpu > gpurun_out/r2z_new_gpu.log 2>&1; tail -8 gpurun_out/r2z_new_gpu.log
