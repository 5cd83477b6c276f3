timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pool" 2>&1 | tail -3
for v in "0 3 8192" "0 2 12288"; do set -- $v
  MG_POOL_VARIANT=$1 MG_POOL_STAGES=$2 MG_POOL_CHUNK=$3 timeout 120 python bench.py --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/pp_v$1_s$2_$3.log 2> gpurun_out/pp_v$1_s$2_$3.err
  python -c "
import json
l=json.loads(open('gpurun_out/pp_v$1_s$2_$3.log').read().strip().splitlines()[-1])
o=l['roofline']['other_kernels']
print('variant $1 stages $2 chunk $3', round(l['ms_per_step']*1e3,1), 'us', round(l['value']), 'img/s lat', round(l['step_latency_ms']*1e3,1), 'pool', round(o['pool_patches_tma_kernel']['ms']*1e3,1), 'us', round(o['pool_patches_tma_kernel']['achieved_gbs']), 'GB/s  unpool', round(l['roofline']['kernel_ms']*1e3,1))
" || tail -3 gpurun_out/pp_v$1_s$2_$3.err
done
