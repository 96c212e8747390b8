mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-400 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --steps 20 --warmup 5 --workload c2 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
