"""How fast is cuFFT itself on the transform sizes the fused kernels use?  (context for the roofline table; not a product path)
    python profiles/cufft_ref.py"""
import json
import torch

dev = torch.device("cuda", 0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n, batch in ((16384, 4096), (16384, 5462), (4096, 16384), (1024, 65536), (2048, 32768)):
    x = torch.view_as_complex(torch.randn(batch, n, 2, device=dev))
    y = torch.empty_like(x)
    ms = timeit(lambda: torch.fft.fft(x, dim=1, out=y))
    print(json.dumps({"cufft_c2c": n, "batch": batch, "ms": round(ms, 4), "gbs_rw": round(2 * x.numel() * 8 / ms / 1e6, 1)}))
for n, batch in ((4096, 32768), (1024, 262144)):
    x = torch.randn(batch, n, device=dev)
    ms = timeit(lambda: torch.fft.rfft(x, dim=1))
    print(json.dumps({"cufft_r2c": n, "batch": batch, "ms": round(ms, 4)}))
