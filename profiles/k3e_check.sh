mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_k3e_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_k3e_tests.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_k3e_v2_R2_W4e6.csv python profiles/run_reduce.py 4000000 3 2 > gpurun_out/r02_k3e_v2_run.log 2>&1
python profiles/run_reduce.py 4000000 6 2 > gpurun_out/r02_k3e_v2_plain.log 2>&1
tail -3 gpurun_out/r02_k3e_tests.log; tail -2 gpurun_out/r02_k3e_v2_plain.log
