mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_final_smoke.log
( time python bench.py --impl reference ) > gpurun_out/r02_final_bench_ref.log 2>&1
( time python bench.py ) > gpurun_out/r02_final_bench_n1.log 2>&1
tail -3 gpurun_out/r02_final_tests.log gpurun_out/r02_final_smoke.log
