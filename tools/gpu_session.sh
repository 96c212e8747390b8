#!/bin/bash
# One complete GPU session (run through `gpurun -- 'bash tools/gpu_session.sh'`): parity suite, smoke, the bench lines of every
# BASELINE config, and the ncu evidence (launch list + `--set full` captures) that tools/summarize_ncu.py turns into
# profiles/*.md.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L | head -1; lscpu | grep "Model name" | head -1
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-300 gpurun_out/bench_c3.json; tail -2 gpurun_out/bench_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
timeout 600 python bench.py --steps 20 --warmup 5 --workload c2 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; cut -c1-300 gpurun_out/bench_c2.json
timeout 900 python bench.py --steps 5 --warmup 3 --workload c4 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-300 gpurun_out/bench_c4.json
timeout 600 python bench.py --steps 60 --warmup 5 --workload c1 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; cut -c1-300 gpurun_out/bench_c1.json
timeout 900 python bench.py --steps 200 --workload c5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; cut -c1-300 gpurun_out/bench_c5.json
if [ "$1" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/b_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_softras -s 3 -c 3 -f -o gpurun_out/prof_softras python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-reference-gpu > gpurun_out/b_ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nmr_ -s 5 -c 5 -f -o gpurun_out/prof_nmr python bench.py --steps 1 --warmup 1 --workload c4 --no-reference-gpu > gpurun_out/b_ncu3.log 2>&1
fi
ls gpurun_out | head -60
