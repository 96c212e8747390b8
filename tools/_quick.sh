mkdir -p gpurun_out
for v in "" _k9in; do
 echo "== c4 lib$v"; B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 600 python bench.py --steps 5 --warmup 3 --workload c4 --no-reference-gpu --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], {k:round(v['avg_ms'],4) for k,v in d['kernels'].items()})"
done
B200R_LIB=$PWD/jrender_b200/lib/libb200raster_k9in.so timeout 600 python -m pytest tests/test_nmr_gpu.py -q -m gpu 2>&1 | tail -2
for wl in c3 c2; do
 for v in "" _bwdfma _allfma; do
  echo "== $wl lib$v"; AB_NO_REF=1 B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so timeout 300 python tools/ab_forward.py $wl 2>&1 | tail -1 | cut -c100-330
 done
done
B200R_LIB=$PWD/jrender_b200/lib/libb200raster_bwdfma.so timeout 900 python -m pytest tests/test_softras_gpu.py tests/test_bake_gpu.py -q -m gpu 2>&1 | tail -8
B200R_LIB=$PWD/jrender_b200/lib/libb200raster_allfma.so timeout 900 python -m pytest tests/test_softras_gpu.py -q -m gpu 2>&1 | tail -15
