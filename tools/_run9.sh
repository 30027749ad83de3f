timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err
cut -c1-300 gpurun_out/r3i_bench.json
