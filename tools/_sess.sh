mkdir -p gpurun_out
B200R_LIB=$PWD/jrender_b200/lib/libb200raster_tma2.so timeout 900 python -m pytest tests/test_softras_gpu.py -q -m gpu -x > gpurun_out/pytest_tma2.log 2>&1; tail -2 gpurun_out/pytest_tma2.log
for wl in c3 c2; do
 for v in "" _tma2 _tma2b; do
  echo "== $wl lib$v"; B200R_LIB=$PWD/jrender_b200/lib/libb200raster$v.so AB_NO_REF=1 timeout 300 python tools/ab_forward.py $wl 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['persistent1'])"
 done
done
