"""How fast can this GPU WRITE a buffer of generate_result's output size (109 MB) that is not in L2?  Plain device fills
(torch fill_ = a store-only kernel, cudaMemsetAsync) timed with CUDA events, L2 flushed in between - the store-only
roofline that generate_result_kernel (31.7 us for the same bytes) is to be read against.  python tools/write_probe.py"""
import json, sys
import torch
n = 111149056 // 8
buf = torch.empty(n, dtype=torch.float64, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
out = {}
for name, fn in (('fill_kernel', lambda: buf.fill_(1.5)), ('memset', lambda: buf.zero_())):
    ts = []
    for _ in range(12):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    out[name] = dict(ms_median=ts[len(ts) // 2], ms_min=ts[0], gbs=n * 8 / (ts[len(ts) // 2] * 1e-3) / 1e9)
print(json.dumps(dict(bytes=n * 8, **out)))
