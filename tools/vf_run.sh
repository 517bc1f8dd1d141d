timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02e_bench_n1.json 2> gpurun_out/r02e_bench_n1.err; echo rc=$?
