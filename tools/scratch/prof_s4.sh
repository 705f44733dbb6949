set -x
python tools/profile_gemm_ln.py 0 > gpurun_out/s4_gemmln_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_ln_split -s 3 -c 1 -o gpurun_out/prof_gemm_ln_split_k512 python tools/profile_gemm_ln.py 0 > gpurun_out/ncu_s4a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ln_split -s 11 -c 1 -o gpurun_out/prof_gemm_ln_split_k2048 python tools/profile_gemm_ln.py 0 > gpurun_out/ncu_s4b.log 2>&1
cat gpurun_out/s4_gemmln_plain.log; tail -2 gpurun_out/ncu_s4a.log gpurun_out/ncu_s4b.log
