cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cholinv.py tests/test_gpu_sequential.py tests/test_gpu_gptq.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/t5.txt 2>&1; grep -v "^$" gpurun_out/t5.txt | grep -E "^(FAILED|ERROR)|passed|failed|Error" | tail -20
timeout 900 python scripts/diag_seq.py 2 > gpurun_out/diag_seq.txt 2>&1; tail -25 gpurun_out/diag_seq.txt
