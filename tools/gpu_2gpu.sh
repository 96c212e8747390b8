mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_c3_2gpu.json 2> gpurun_out/bench_c3_2gpu.err; cut -c1-300 gpurun_out/bench_c3_2gpu.json; tail -3 gpurun_out/bench_c3_2gpu.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('allgather_images_ms'), d.get('gather_in_step'))"
