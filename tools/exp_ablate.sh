#!/bin/bash
# Timing ablations of the run kernel (gpurun_exp/libni_*.so built with -DNI_ABLATE=k, see ni_core.cuh): prints SM
# cycles per warp-attempt per scheduler for the product library and every variant.
# usage: gpurun -- 'bash tools/exp_ablate.sh [bench args]'
mkdir -p gpurun_out
LIB=numba_integrators_b200/libni_b200.so
cp $LIB /tmp/libni_product.so
ARGS="${@:---traj 1000000 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e}"
run() {
  python bench.py $ARGS > gpurun_out/abl_$1.log 2>&1
  python - "$1" <<'PY'
import json,sys
tag=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/abl_{tag}.log').read().strip().splitlines()[-1]); r=d['roofline']
    cyc=r['kernel_ms']*1e-3*d['clocks']['sm_mhz']*1e6*148*4/(r['attempts_per_launch']/32)
    print(f"== {tag}: kernel {r['kernel_ms']:.3f} ms, attempts {r['attempts_per_launch']}, {cyc:.1f} cycles per warp-attempt per scheduler, sm {d['clocks']['sm_mhz']} MHz")
except Exception as e:
    print('==',tag,'failed',e)
PY
}
run product
for f in gpurun_exp/libni_*.so; do
  [ -e "$f" ] || continue
  tag=$(basename $f .so); tag=${tag#libni_}
  cp $f $LIB
  run $tag
done
cp /tmp/libni_product.so $LIB
