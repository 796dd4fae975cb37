mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:"k_window_rows|k_xs_compose|k_xs_walk|k_window_reduce" -s 5 -c 5 -f -o gpurun_out/r02_ncu_staged_R2 python profiles/run_reduce.py 4000000 3 2 > gpurun_out/r02_ncu_staged.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"k_window_fused" -s 3 -c 1 -f -o gpurun_out/r02_ncu_fused python profiles/run_reduce_native.py 4000000 5 > gpurun_out/r02_ncu_fused.log 2>&1
ncu --set full --clock-control none -k regex:"k_commit|k_stamp_end|k_mirror" -s 6 -c 3 -f -o gpurun_out/r02_ncu_steppath python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_ncu_steppath.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_n1.csv python bench.py --steps 2 --warmup 3 --no-overhead --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench_n1.log 2>&1
ls -la gpurun_out/*.ncu-rep
