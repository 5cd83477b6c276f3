"""GPU micro-benchmark of the patch-pool kernel variants (MG_POOL_VARIANT is read once per process)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
import mingraph_unet_b200 as mg
x = torch.randn(16, 20, 512, 512, device="cuda").bfloat16()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
ref = None
ts = []
for i in range(30):
    flush.zero_()
    torch.cuda._sleep(200000)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); y = mg.ops.pool_patches(x, 16); e1.record(); torch.cuda.synchronize()
    if i >= 5: ts.append(e0.elapsed_time(e1))
ts.sort()
print("%%.1f us median, %%.1f us min, %%.0f GB/s, checksum %%.6f" %% (1e3*ts[len(ts)//2], 1e3*ts[0], x.numel()*2/ts[len(ts)//2]/1e6, float(y.float().sum())))
''' % ROOT
for v in sys.argv[1:] or ["0", "1", "2", "5", "8"]:
    env = dict(os.environ, MG_POOL_VARIANT=v)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print("variant", v, out.stdout.strip(), out.stderr.strip()[-300:])
