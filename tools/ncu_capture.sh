set -x
cd $GRAFT_REPO_ROOT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"conv_umma|conv_rowwin|pool_kernel|gap_dense" -s 23 -c 23 -o gpurun_out/r02b_resnet18_kernels -f python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-extra > gpurun_out/r02b_ncu_full.log 2>&1
timeout 600 ncu --graph-profiling node --metrics gpu__time_duration.sum --clock-control none -s 66 -c 66 --csv --log-file gpurun_out/r02b_launches_bench_resnet18.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02b_launches.log 2>&1
SNNB_UMMA_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-extra > gpurun_out/r02b_trace_raw.txt 2>&1
ls -la gpurun_out | tail -8
