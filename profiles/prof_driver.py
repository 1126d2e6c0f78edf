"""Short driver for ncu captures: one warm-up + `reps` launches of each headline kernel at bench sizes.
    ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 -o gpurun_out/<name> python profiles/prof_driver.py <what>
what: conv | welch_c | welch_r | spectro | resample | fir
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dspb200 import _lib  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "conv"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
log2n = int(sys.argv[3]) if len(sys.argv) > 3 else 26
nfft = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
n = 1 << log2n
win = bench.hanning64(4096)
norm2 = float(np.sum(win * win))
if what == "conv":
    x = torch.view_as_complex(torch.randn(n, 2, device=dev))
    y = torch.empty(n + bench.NV - 1, dtype=torch.complex64, device=dev)
    plan = _lib.OsPlan(bench.make_taps(), nfft)
    for _ in range(reps):
        plan.exec_dev(x.data_ptr(), n, 1, y.data_ptr(), y.numel(), 0)
elif what == "welch_c":
    x = torch.view_as_complex(torch.randn(n, 2, device=dev))
    p = torch.empty(4096, dtype=torch.float32, device=dev)
    plan = _lib.SpecPlan(np.complex64, 4096, 2048, 4096, False, win)
    for _ in range(reps):
        plan.welch_dev(x.data_ptr(), n, plan.nsegments(n) * norm2, p.data_ptr(), 0)
elif what == "welch_r":
    x = torch.randn(n, device=dev)
    p = torch.empty(2049, dtype=torch.float32, device=dev)
    plan = _lib.SpecPlan(np.float32, 4096, 2048, 4096, True, win)
    for _ in range(reps):
        plan.welch_dev(x.data_ptr(), n, plan.nsegments(n) * norm2, p.data_ptr(), 0)
elif what == "spectro":
    nchan, length = 64, 1 << 22
    x = torch.randn(nchan * length, device=dev)
    plan = _lib.SpecPlan(np.float32, 1024, 768, 1024, True, None)
    k = plan.nsegments(length)
    out = torch.empty(513 * k * nchan, dtype=torch.float32, device=dev)
    for _ in range(reps):
        plan.stft_dev(x.data_ptr(), length, nchan, 1024.0, True, out.data_ptr(), 0)
elif what == "resample":
    import dspb200
    from fractions import Fraction
    h = dspb200.resample_filter(Fraction(3, 2)).astype(np.float32)
    x = torch.view_as_complex(torch.randn(n, 2, device=dev))
    nout = 3 * n // 2
    y = torch.empty(nout, dtype=torch.complex64, device=dev)
    plan = _lib.ResamplePlan(np.complex64, h, 3, 2)
    n0, phi0 = dspb200.resample_phase(h.size, Fraction(3, 2))
    for _ in range(reps):
        plan.exec_dev(x.data_ptr(), n, 1, n0, phi0, y.data_ptr(), nout, 0)
elif what == "fir":
    nn = np.arange(257) - 128
    b = (0.5 * np.sinc(0.5 * nn) * np.hamming(257)).astype(np.float32)
    x = torch.randn(1 << 20, device=dev)
    y = torch.empty_like(x)
    plan = _lib.FirPlan(b)
    for _ in range(reps):
        plan.exec_dev(x.data_ptr(), 1 << 20, 1, y.data_ptr(), 0)
torch.cuda.synchronize()
print("done", what, _lib.launch_count())
