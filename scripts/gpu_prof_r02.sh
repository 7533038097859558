cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# (1) launch list of the bench command (main step only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
python scripts/ncu_agg.py gpurun_out/r02_launches_bench.csv | head -8
# (2) full capture of the bench kernel
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_strided_multi -c 1 -f -o gpurun_out/r02_prof_strided_multi python scripts/prof_targets.py quant > gpurun_out/r02_ncu_q.log 2>&1; tail -2 gpurun_out/r02_ncu_q.log
# (3) GPTQ solve: launch list + full captures of the three solver kernels
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_gptq.csv python scripts/prof_targets.py gptq > gpurun_out/r02_ncu_g.log 2>&1
python scripts/ncu_agg.py gpurun_out/r02_launches_gptq.csv | head -10
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_gptq_block2|k_chol_leaf" -s 40 -c 4 -f -o gpurun_out/r02_prof_gptq_kernels python scripts/prof_targets.py gptq > gpurun_out/r02_ncu_g2.log 2>&1; tail -2 gpurun_out/r02_ncu_g2.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gemm_split3 -s 300 -c 3 -f -o gpurun_out/r02_prof_split3 python scripts/prof_targets.py gptq > gpurun_out/r02_ncu_g3.log 2>&1; tail -2 gpurun_out/r02_ncu_g3.log
# (4) packed GEMMs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_rb|k_gemm_mx8" -c 4 -f -o gpurun_out/r02_prof_gemm python scripts/prof_targets.py gemm > gpurun_out/r02_ncu_gemm.log 2>&1; tail -2 gpurun_out/r02_ncu_gemm.log
ls -la gpurun_out/*.ncu-rep | tail -6
