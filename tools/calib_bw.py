"""HBM calibration: what do plain torch kernels reach on read-only / write-only / copy traffic of the bench's sizes?"""
import torch
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def timeit(fn, n=20):
    ts = []
    for i in range(n + 3):
        flush.zero_(); torch.cuda._sleep(200000)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= 3: ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for mb in (168, 537, 2148):
    n = mb * 1000 * 1000 // 2
    x = torch.randn(n, device="cuda").bfloat16(); y = torch.empty_like(x)
    t = timeit(lambda: torch.sum(x));           print(f"{mb} MB  read  (torch.sum bf16)   {t:7.1f} us  {n*2/t/1e3:7.0f} GB/s")
    t = timeit(lambda: y.fill_(1.0));           print(f"{mb} MB  write (fill_)            {t:7.1f} us  {n*2/t/1e3:7.0f} GB/s")
    t = timeit(lambda: y.copy_(x));             print(f"{mb} MB  copy  (read+write)       {t:7.1f} us  {2*n*2/t/1e3:7.0f} GB/s")
    xf = x.view(torch.float32)
    t = timeit(lambda: torch.sum(xf));          print(f"{mb} MB  read  (torch.sum f32)    {t:7.1f} us  {n*2/t/1e3:7.0f} GB/s")
