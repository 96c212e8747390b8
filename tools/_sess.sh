mkdir -p gpurun_out
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/check_peer_broadcast.py > gpurun_out/peer_check.log 2>&1; grep -E "PEER|Error|error" gpurun_out/peer_check.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_c3_${N}gpu.json 2> gpurun_out/bench_c3_${N}gpu.err; grep -v "CudaIPCTypes\|OMP_NUM\|^\*\*\*" gpurun_out/bench_c3_${N}gpu.err | tail -3
python -c "
import json; d=json.loads(open('gpurun_out/bench_c3_${N}gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d.get('allgather_images_ms')); g=d.get('gather_in_step'); g.pop('note',None); print(json.dumps(g))"
