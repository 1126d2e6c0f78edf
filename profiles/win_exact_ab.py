"""A/B of the Float32 window product (spectral.cu, win_mul): the default `fma(x, wh, x*wl)` on a hi/lo float pair against
the -DDSP_WIN_EXACT=1 build (sample widened, one DMUL, rounded back -- the reference's own rounding, src/periodograms.jl:66).
Run once per build:   [DSPB200_LIB=dsp.jl_b200/libdspb200_winexact.so] python profiles/win_exact_ab.py
Prints how many windowed samples differ from the oracle's reference-rounded product, and the times of the windowed kernels."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dspb200  # noqa: E402
from dspb200 import _lib  # noqa: E402
from oracle import periodograms as op  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
tag = os.path.basename(_lib.LIB_PATH)
rng = np.random.default_rng(66)
win = bench.hanning64(4096)
for dt in (np.float32, np.complex64):
    s = rng.standard_normal(1 << 22).astype(np.float32)
    if dt is np.complex64:
        s = (s + 1j * rng.standard_normal(1 << 22).astype(np.float32)).astype(np.complex64)
    got = dspb200.arraysplit(s, 4096, 2048, window=win)
    want = op.arraysplit(s, 4096, 2048, 4096, win)
    got = np.asarray(got).reshape(want.shape) if np.asarray(got).shape != want.shape else np.asarray(got)
    diff = int(np.count_nonzero(got.view(np.float32) != want.view(np.float32)))
    print(json.dumps({"lib": tag, "check": f"arraysplit {np.dtype(dt).name} 2^22 samples, hanning(4096), 50 %", "values": int(want.size * (2 if dt is np.complex64 else 1)),
                      "differ_from_reference_rounding": diff}), flush=True)


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
norm2 = float(np.sum(win * win))
xr = torch.randn(n, device=dev)
xc = torch.view_as_complex(torch.randn(n, 2, device=dev))
sp3 = _lib.SpecPlan(np.float32, 4096, 2048, 4096, True, win)
p3 = torch.empty(2049, device=dev)
k3 = sp3.nsegments(n)
rows = [("C3 welch_pgram 2^26 F32 4096/50 %/hanning", timeit(lambda: sp3.welch_dev(xr.data_ptr(), n, k3 * norm2, p3.data_ptr(), 0)))]
sp3c = _lib.SpecPlan(np.complex64, 4096, 2048, 4096, False, win)
p3c = torch.empty(4096, device=dev)
rows.append(("welch_pgram 2^26 CF32 4096/50 %/hanning", timeit(lambda: sp3c.welch_dev(xc.data_ptr(), n, k3 * norm2, p3c.data_ptr(), 0))))
del xc
nchan, length = 64, 1 << 22
x4 = torch.randn(nchan * length, device=dev)
w1k = bench.hanning64(1024)
sp4 = _lib.SpecPlan(np.float32, 1024, 768, 1024, True, w1k)
k4 = sp4.nsegments(length)
o4 = torch.empty(513 * k4 * nchan, device=dev)
rows.append(("C4 spectrogram 64 x 2^22 F32 1024/75 %/hanning", timeit(lambda: sp4.stft_dev(x4.data_ptr(), length, nchan, float(np.sum(w1k * w1k)), True, o4.data_ptr(), 0), reps=10)))
sp5 = _lib.SpecPlan(np.float32, 4096, 2048, 4096, True, win)
k5 = sp5.nsegments(length)
o5 = torch.empty(2049 * k5 * 8, device=dev)
rows.append(("spectrogram 8 x 2^22 F32 4096/50 %/hanning", timeit(lambda: sp5.stft_dev(x4.data_ptr(), length, 8, norm2, True, o5.data_ptr(), 0), reps=10)))
for name, ms in rows:
    print(json.dumps({"lib": tag, "config": name, "ms": round(ms, 4)}), flush=True)
