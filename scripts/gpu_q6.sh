cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_quant.py tests/test_gpu_pack.py tests/test_gpu_packed_gemm.py tests/test_gpu_mx8.py tests/test_ref_native.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python scripts/gpu_sweep.py 2>&1 | grep -E "axis1|pack"
