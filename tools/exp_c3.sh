#!/bin/bash
# A/B of run-kernel variants on the default bench (C3): the product library, then every gpurun_exp/libni_*.so copied
# over it (the GPU box is a scratch copy).  usage: gpurun -- 'bash tools/exp_c3.sh [bench args]'
mkdir -p gpurun_out
LIB=numba_integrators_b200/libni_b200.so
cp $LIB /tmp/libni_product.so
ARGS="${@:---steps 3 --warmup 3 --no-cpu-baseline --no-e2e}"
run() {
  python bench.py $ARGS > gpurun_out/exp_$1.log 2>&1
  echo "== $1 rc=$? $(grep -o '"value": [0-9.e+]*' gpurun_out/exp_$1.log | head -1) $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/exp_$1.log) $(grep -o '"frac": [0-9.]*' gpurun_out/exp_$1.log | head -1)"
}
run product
for f in gpurun_exp/libni_*.so; do
  [ -e "$f" ] || continue
  tag=$(basename $f .so); tag=${tag#libni_}
  cp $f $LIB
  run $tag
done
cp /tmp/libni_product.so $LIB
