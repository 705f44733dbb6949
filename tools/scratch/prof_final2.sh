set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r02g_bench_n1.json 2> gpurun_out/r02g_bench_n1.err
python tools/kernel_table.py > gpurun_out/r02g_ktable.txt 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-strong > gpurun_out/r02g_bench_short.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02g_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-gpu --no-strong > gpurun_out/r02g_ncu_bench.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02g_smoke.log 2>&1; tail -1 gpurun_out/r02g_smoke.log
head -c 400 gpurun_out/r02g_bench_n1.json; grep -v "^/opt\|_warn_once" gpurun_out/r02g_ktable.txt | head -8
