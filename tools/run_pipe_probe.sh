for c in 1 2 4 8; do
  MG_BLOCK_CLUSTER=$c python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "block or fused" 2>&1 | tail -1
  for cfg in "1 2" "1 3" "2 2"; do set -- $cfg
    MG_BLOCK_CLUSTER=$c python bench.py --steps 300 --warmup 10 --shards $1 --pipeline-depth $2 --no-cpu-baseline > gpurun_out/pc_c${c}_s$1_d$2.log 2> gpurun_out/pc_c${c}_s$1_d$2.err
    python -c "
import json
l=json.loads(open('gpurun_out/pc_c${c}_s$1_d$2.log').read().strip().splitlines()[-1])
print('cluster $c shards $1 depth $2', round(l['ms_per_step']*1e3,1), 'us', round(l['value']), 'img/s lat', round(l['step_latency_ms']*1e3,1), 'block', round(l['roofline']['other_kernels']['block_forward_kernel']['ms']*1e3,1))
" || tail -3 gpurun_out/pc_c${c}_s$1_d$2.err
  done
done
