"""A/B of the two multi-phase resample kernels (DSPB200_RS_MP2=0: resample_mp_kernel, default: the pipelined
resample_mp2_kernel) on BASELINE config 5 and a few other ratios; run once per setting.   python profiles/resample_ab.py"""
import json
import os
import sys
from fractions import Fraction

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dspb200  # noqa: E402
from dspb200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
mode = "mp (DSPB200_RS_MP2=0)" if os.environ.get("DSPB200_RS_MP2", "1")[0] == "0" else "mp2 (pipelined, default)"
if os.environ.get("DSPB200_LIB"):
    mode += " [" + os.path.basename(os.environ["DSPB200_LIB"]) + "]"


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


n = 1 << 26
xc = torch.view_as_complex(torch.randn(n, 2, device=dev))
xr = torch.randn(n, device=dev)
for (i, d) in ((3, 2), (2, 1), (3, 1), (4, 3), (2, 3), (3, 4)):
    rate = Fraction(i, d)
    h = dspb200.resample_filter(rate)
    n0, phi0 = dspb200.resample_phase(h.size, rate)
    nout = n * i // d
    for name, xin, hh, odt in (("CF32, F32 taps", xc, h.astype(np.float32), torch.complex64), ("F32, F32 taps", xr, h.astype(np.float32), torch.float32),
                               ("CF32, F64 taps", xc, h, torch.complex128)):
        if odt == torch.complex128 and (i, d) != (3, 2):
            continue
        y = torch.empty(nout, dtype=odt, device=dev)
        plan = _lib.ResamplePlan(np.complex64 if xin.is_complex() else np.float32, hh, i, d)
        ms = timeit(lambda: plan.exec_dev(xin.data_ptr(), n, 1, n0, phi0, y.data_ptr(), nout, 0), reps=10)
        print(json.dumps({"kernel": mode, "config": f"resample {i}//{d} 2^26 {name} ({h.size} taps)", "ms": round(ms, 4)}), flush=True)
        del y, plan
