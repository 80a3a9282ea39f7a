# Round-end measurements on a one-GPU box (gpurun -- bash tools/final_profiles.sh); results under gpurun_out/, copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.txt
timeout 600 python bench.py > gpurun_out/r2_bench_line.json 2> gpurun_out/r2_bench_line.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_line.json 2> gpurun_out/r2_bench_reference_line.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --fields 128 > gpurun_out/ncu_bench.log 2>&1
timeout 200 python tools/time_c1.py 2>&1 | tail -1 > gpurun_out/r2_c1.json
timeout 200 python tools/time_1d.py 26 > gpurun_out/r2_1d_2e26.json 2>&1
tail -3 gpurun_out/r2_pytest_gpu.txt; tail -c 600 gpurun_out/r2_bench_line.json; tail -c 300 gpurun_out/r2_bench_reference_line.json
