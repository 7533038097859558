cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pack.py tests/test_gpu_mx4.py tests/test_gpu_mx8.py tests/test_gpu_packed_gemm.py -x -q 2>&1 | tail -15
echo "--- epilogue warps A/B (default 8 / MSQ_MX8T_EPI4=1)"
timeout 200 python scripts/cta2_test.py 2>&1 | tail -6 | sed 's/identical.*| 1-CTA/|/'
MSQ_MX8T_EPI4=1 timeout 200 python scripts/cta2_test.py 2>&1 | tail -6 | sed 's/identical.*| 1-CTA/|/'
