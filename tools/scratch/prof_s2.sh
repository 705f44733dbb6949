set -x
python tools/profile_step.py > gpurun_out/step_plain_s2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_s2_metrics.csv python tools/profile_step.py > gpurun_out/ncu_step_s2.log 2>&1
python tools/profile_gemm.py > gpurun_out/gemm_plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05_kernel -s 4 -c 1 -o gpurun_out/prof_gemm_ffn1_s2 python tools/profile_gemm.py > gpurun_out/ncu_gemm_s2.log 2>&1
python tools/profile_ops.py attn1 dwconv > gpurun_out/ops_plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"attention_tc_kernel|dwconv_ln_silu_tma" -s 6 -c 2 -o gpurun_out/prof_ops_s2 python tools/profile_ops.py attn1 dwconv > gpurun_out/ncu_ops_s2.log 2>&1
cat gpurun_out/step_plain_s2.log gpurun_out/gemm_plain_s2.log gpurun_out/ops_plain_s2.log
tail -2 gpurun_out/ncu_step_s2.log gpurun_out/ncu_gemm_s2.log gpurun_out/ncu_ops_s2.log
