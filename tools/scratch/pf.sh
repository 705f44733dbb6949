export CHUNKFORMER_B200_LIB=chunkformer_b200/csrc/libchunkformer_b200_ablation.so
for pf in 0 4 8 16 32 64; do echo "== prefetch $pf"; CF_LN_PREFETCH=$pf timeout 120 python tools/ablate_gemm_ln.py 0 63; done
