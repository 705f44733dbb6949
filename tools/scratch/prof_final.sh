set -x
# 1. bench (full line) and reference arm
python bench.py --steps 5 --warmup 3 > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/r02f_bench_ref.err
# 2. kernel table (CUPTI)
python tools/kernel_table.py > gpurun_out/r02f_ktable.txt 2>&1
# 3. launch list of the bench command under ncu (after the plain run above exited 0)
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-strong > gpurun_out/r02f_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-strong > gpurun_out/r02f_ncu_bench.log 2>&1
# 4. full captures: FFN w_1 (pair GEMM), residual GEMM + LayerNorm (K = 512, K = 2048), attention
python tools/profile_gemm.py > gpurun_out/r02f_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm2_tcgen05_kernel -s 4 -c 1 -o gpurun_out/r02f_prof_gemm_ffn1 python tools/profile_gemm.py > gpurun_out/r02f_ncu_gemm.log 2>&1
python tools/profile_gemm_ln.py 0 > gpurun_out/r02f_gemm_ln_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_ln_split -s 3 -c 1 -o gpurun_out/r02f_prof_gemm_ln_k512 python tools/profile_gemm_ln.py 0 > gpurun_out/r02f_ncu_ln1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_ln_split -s 11 -c 1 -o gpurun_out/r02f_prof_gemm_ln_k2048 python tools/profile_gemm_ln.py 0 > gpurun_out/r02f_ncu_ln2.log 2>&1
python tools/profile_ops.py attn1 > gpurun_out/r02f_attn_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 4 -c 1 -o gpurun_out/r02f_prof_attn python tools/profile_ops.py attn1 > gpurun_out/r02f_ncu_attn.log 2>&1
cat gpurun_out/r02f_gemm_plain.log gpurun_out/r02f_gemm_ln_plain.log gpurun_out/r02f_attn_plain.log
head -c 600 gpurun_out/r02f_bench_n1.json
