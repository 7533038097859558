cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python scripts/llama7b_e2e.py --layers 32 > gpurun_out/llama7b_e2e.json 2> gpurun_out/llama7b_e2e.err; echo rc=$?
tail -3 gpurun_out/llama7b_e2e.json; tail -5 gpurun_out/llama7b_e2e.err
timeout 600 python scripts/llama7b_e2e.py --layers 32 --nearest > gpurun_out/llama7b_rtn.json 2>> gpurun_out/llama7b_e2e.err; tail -2 gpurun_out/llama7b_rtn.json
