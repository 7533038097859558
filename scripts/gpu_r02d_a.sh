cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python scripts/pack_cvt_adv.py > gpurun_out/r02d_pack_cvt_adv.txt 2>&1; echo "adv rc $?"; tail -15 gpurun_out/r02d_pack_cvt_adv.txt
timeout 300 python scripts/pack_ab.py > gpurun_out/r02d_pack_ab.txt 2>&1; echo "ab rc $?"; tail -8 gpurun_out/r02d_pack_ab.txt
MSQ_PACK_CVT=0 timeout 300 python scripts/pack_ab.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_pack.py -x -q 2>&1 | tail -3
echo "--- epilogue warps A/B (default 8 / MSQ_MX8T_EPI4=1)"
timeout 200 python scripts/cta2_test.py 2>&1 | tail -6 | cut -c1-60,150-260
MSQ_MX8T_EPI4=1 timeout 200 python scripts/cta2_test.py 2>&1 | tail -6 | cut -c1-60,150-260
