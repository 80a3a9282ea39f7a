set -x
timeout 400 python bench.py --steps 20 --warmup 4 2>&1 | tail -1 > gpurun_out/r1_bench_line.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r1_bench_reference_line.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
PRECISION=fp64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fb_sweep -s 2 -c 2 -o gpurun_out/prof_final64 -f python tools/run_fp32_once.py > gpurun_out/ncu64.log 2>&1
PRECISION=fp32 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fb_sweep -s 2 -c 2 -o gpurun_out/prof_final32 -f python tools/run_fp32_once.py > gpurun_out/ncu32.log 2>&1
ls -la gpurun_out | tail -8
