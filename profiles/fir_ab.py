"""A/B of the two time-domain FIR kernels (DSPB200_FIR_TILE=0: fir_td_kernel, default: the register-tiled
fir_tile_kernel) on BASELINE config 1 and wider shapes; also checks that both produce the same bits.  Run once per setting.
    python profiles/fir_ab.py"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dspb200  # noqa: E402,F401
from dspb200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
mode = "fir_td_kernel (DSPB200_FIR_TILE=0)" if os.environ.get("DSPB200_FIR_TILE", "1")[0] == "0" else "fir_tile_kernel v2 (default)"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps):
        fn()
    b.record(st)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


rng = np.random.default_rng(5)
torch.manual_seed(5)
cases = [("C1: 257 taps, 2^20 F32", np.float32, 257, 1 << 20, 1),
         ("257 taps, 2^26 F32 (64 columns of 2^20)", np.float32, 257, 1 << 20, 64),
         ("257 taps, 2^24 CF32", np.complex64, 257, 1 << 24, 1),
         ("257 taps, 2^24 F64", np.float64, 257, 1 << 24, 1),
         ("257 taps, 2^23 CF64", np.complex128, 257, 1 << 23, 1),
         ("66 taps, 2^24 F32", np.float32, 66, 1 << 24, 1),
         ("1500 taps, 2^22 F32", np.float32, 1500, 1 << 22, 1)]
tdt = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64, np.complex128: torch.complex128}
for name, dt, nb, nx, ncols in cases:
    cplx = np.dtype(dt).kind == "c"
    b = rng.standard_normal(nb) + (1j * rng.standard_normal(nb) if cplx else 0)
    b = b.astype(dt)
    x = torch.randn(nx * ncols * (2 if cplx else 1), device=dev, dtype=tdt[dt].to_real() if cplx else tdt[dt])
    if cplx:
        x = torch.view_as_complex(x.view(-1, 2))
    y = torch.empty_like(x)
    plan = _lib.FirPlan(b)
    ms = timeit(lambda: plan.exec_dev(x.data_ptr(), nx, ncols, y.data_ptr(), 0))
    flops = 2.0 * (4 if cplx else 1) * nb * nx * ncols
    # order-independent fingerprint of the output bits (the two kernels run the same fma chain: equal fingerprints)
    bits = y.view(torch.uint8).to(torch.int64)
    fp = int((bits * (torch.arange(bits.numel(), device=dev) % 251 + 1)).sum().item())
    print(json.dumps({"kernel": mode, "config": name, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 2),
                      "gsamples_s": round(nx * ncols / ms / 1e6, 2), "output_fingerprint": fp}), flush=True)
    del x, y, plan, bits
    torch.cuda.empty_cache()
