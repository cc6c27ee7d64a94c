#!/bin/bash
# multi-GPU validation + numbers: gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN tools/multi_gpu_check.py > gpurun_out/r2_multi_check_$N.log 2>&1; echo "multi_gpu_check rc=$?"; grep -E "gather ==|Error|error" gpurun_out/r2_multi_check_$N.log | tail -5
NCCL_DEBUG=INFO $RUN bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2>&1; echo "bench rc=$?"
grep -m3 "NVLS\|nranks" gpurun_out/r2_bench_${N}gpu.log | cut -c1-200
grep "^{\"metric" gpurun_out/r2_bench_${N}gpu.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g  ms %.2f  gather %.3f ms  incl %.4g  GB/s/rank %.0f'%(d['value'], d['ms_per_step'], d['final_gather_ms'], d['value_incl_gather'], d['gather_GBps_per_rank']))
s=d['strong']; print('strong:', {k:(round(v,4) if isinstance(v,float) else v) for k,v in s.items() if k!='note'})
print('e2e %.4g frac %.4f fast %.4f'%(d['e2e']['value'], d['roofline']['frac'], d['roofline']['fast']['frac']))
print('fused', d.get('fused_gather')); print('ordered', d.get('cost_ordered'))"
