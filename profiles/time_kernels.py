"""CUDA-event timing of every BASELINE config at full size (device-resident), one JSON line per config.
    python profiles/time_kernels.py [reps]"""
import json
import os
import sys
from fractions import Fraction

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import dspb200  # noqa: E402
from dspb200 import _lib  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
peak, _ = bench.measured_peak_gbs()
st = torch.cuda.current_stream()


def timeit(fn):
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


def report(name, ms, samples, bytes_):
    print(json.dumps({"config": name, "ms": round(ms, 4), "gsamples_s": round(samples / ms / 1e6, 2),
                      "algorithmic_gbs": round(bytes_ / ms / 1e6, 1), "frac_of_measured_hbm": round(bytes_ / ms / 1e6 / peak, 4)}))


# C1: filt(b,1,x) 257-tap on 2^20 F32 (time domain) and the fftfilt equivalent
n1 = 1 << 20
nn = np.arange(257) - 128
b = (0.5 * np.sinc(0.5 * nn) * np.hamming(257)).astype(np.float32)
x1 = torch.randn(n1, device=dev)
y1 = torch.empty_like(x1)
fir = _lib.FirPlan(b)
report("C1 filt(b,1,x) 257-tap 2^20 F32 (fir_tile_kernel)", timeit(lambda: fir.exec_dev(x1.data_ptr(), n1, 1, y1.data_ptr(), 0)), n1, 8 * n1)
osr = _lib.OsPlan(b, 0)
report(f"C1' fftfilt 257-tap 2^20 F32 (fused nfft={osr.nfft})", timeit(lambda: osr.exec_dev(x1.data_ptr(), n1, 1, y1.data_ptr(), n1, 0)), n1, 8 * n1)
n = 1 << 26
xr = torch.randn(n, device=dev)
yr = torch.empty_like(xr)
report(f"fftfilt 257-tap 2^26 F32 (fused nfft={osr.nfft})", timeit(lambda: osr.exec_dev(xr.data_ptr(), n, 1, yr.data_ptr(), n, 0)), n, 8 * n)
taps_r = np.real(bench.make_taps()).astype(np.float32)
osr2 = _lib.OsPlan(taps_r, 0)
report(f"fftfilt 4097-tap 2^26 F32 (fused nfft={osr2.nfft})", timeit(lambda: osr2.exec_dev(xr.data_ptr(), n, 1, yr.data_ptr(), n, 0)), n, 8 * n)

# C2: conv 4097-tap on 2^26 CF32
xc = torch.view_as_complex(torch.randn(n, 2, device=dev))
yc = torch.empty(n + bench.NV - 1, dtype=torch.complex64, device=dev)
osc = _lib.OsPlan(bench.make_taps(), 0)
report(f"C2 conv 4097-tap 2^26 CF32 (fused nfft={osc.nfft})", timeit(lambda: osc.exec_dev(xc.data_ptr(), n, 1, yc.data_ptr(), yc.numel(), 0)), n, 16 * n)
osg = _lib.OsPlan(bench.make_taps(), 65536)
report("C2 conv 4097-tap 2^26 CF32 (cuFFT path, nfft=65536 as the reference picks)", timeit(lambda: osg.exec_dev(xc.data_ptr(), n, 1, yc.data_ptr(), yc.numel(), 0)), n, 16 * n)

# C3: welch 2^26 F32
win = bench.hanning64(4096)
norm2 = float(np.sum(win * win))
sp3 = _lib.SpecPlan(np.float32, 4096, 2048, 4096, True, win)
p3 = torch.empty(2049, device=dev)
k3 = sp3.nsegments(n)
report("C3 welch_pgram 2^26 F32 nfft=4096 50% hanning", timeit(lambda: sp3.welch_dev(xr.data_ptr(), n, k3 * norm2, p3.data_ptr(), 0)), n, 4 * n)
sp3c = _lib.SpecPlan(np.complex64, 4096, 2048, 4096, False, win)
p3c = torch.empty(4096, device=dev)
report("welch_pgram 2^26 CF32 nfft=4096 50% hanning (two-sided)", timeit(lambda: sp3c.welch_dev(xc.data_ptr(), n, k3 * norm2, p3c.data_ptr(), 0)), n, 8 * n)

# C4: spectrogram 64 x 2^22 F32, nfft=1024, 75 %
nchan, length = 64, 1 << 22
x4 = torch.randn(nchan * length, device=dev)
sp4 = _lib.SpecPlan(np.float32, 1024, 768, 1024, True, None)
k4 = sp4.nsegments(length)
o4 = torch.empty(513 * k4 * nchan, device=dev)
report("C4 spectrogram 64ch x 2^22 F32 nfft=1024 75% (rect)", timeit(lambda: sp4.stft_dev(x4.data_ptr(), length, nchan, 1024.0, True, o4.data_ptr(), 0)),
       nchan * length, 4 * nchan * length + 4 * 513 * k4 * nchan)
del x4, o4

# C5: resample 3//2 2^26 CF32, F32 taps -> CF32, and F64 default taps -> CF64
h = dspb200.resample_filter(Fraction(3, 2))
n0, phi0 = dspb200.resample_phase(h.size, Fraction(3, 2))
nout = 3 * n // 2
y5 = torch.empty(nout, dtype=torch.complex64, device=dev)
rs = _lib.ResamplePlan(np.complex64, h.astype(np.float32), 3, 2)
report("C5 resample 3//2 2^26 CF32, F32 taps -> CF32", timeit(lambda: rs.exec_dev(xc.data_ptr(), n, 1, n0, phi0, y5.data_ptr(), nout, 0)), n, 20 * n)
del y5
y5d = torch.empty(nout, dtype=torch.complex128, device=dev)
rsd = _lib.ResamplePlan(np.complex64, h, 3, 2)
report("C5 resample 3//2 2^26 CF32, default F64 taps -> CF64", timeit(lambda: rsd.exec_dev(xc.data_ptr(), n, 1, n0, phi0, y5d.data_ptr(), nout, 0)), n, 8 * n + 24 * n)
del y5d

# Float64 / ComplexF64 rows (north_star quotes a 1e-12 tolerance for them): the fused transforms reach 8192 points in double
# precision, so the 4097-tap filter runs with nfft = 8192 (L = 4096)
n64 = 1 << 25
xd = torch.view_as_complex(torch.randn(n64, 2, device=dev, dtype=torch.float64))
yd = torch.empty(n64 + bench.NV - 1, dtype=torch.complex128, device=dev)
osd = _lib.OsPlan(bench.make_taps().astype(np.complex128), 0)
report(f"conv 4097-tap 2^25 CF64 (nfft={osd.nfft}, fused={osd.fused})", timeit(lambda: osd.exec_dev(xd.data_ptr(), n64, 1, yd.data_ptr(), yd.numel(), 0)), n64, 32 * n64)
del yd
xrd = torch.randn(n64, device=dev, dtype=torch.float64)
yrd = torch.empty(n64, dtype=torch.float64, device=dev)
osrd = _lib.OsPlan(np.real(bench.make_taps()).astype(np.float64), 0)
report(f"fftfilt 4097-tap 2^25 F64 (nfft={osrd.nfft}, fused={osrd.fused})", timeit(lambda: osrd.exec_dev(xrd.data_ptr(), n64, 1, yrd.data_ptr(), n64, 0)), n64, 16 * n64)
del yrd
sp3d = _lib.SpecPlan(np.float64, 4096, 2048, 4096, True, win)
p3d = torch.empty(2049, device=dev, dtype=torch.float64)
k3d = sp3d.nsegments(n64)
report("welch_pgram 2^25 F64 nfft=4096 50% hanning", timeit(lambda: sp3d.welch_dev(xrd.data_ptr(), n64, k3d * norm2, p3d.data_ptr(), 0)), n64, 8 * n64)
sp3z = _lib.SpecPlan(np.complex128, 4096, 2048, 4096, False, win)
p3z = torch.empty(4096, device=dev, dtype=torch.float64)
report("welch_pgram 2^25 CF64 nfft=4096 50% hanning (two-sided)", timeit(lambda: sp3z.welch_dev(xd.data_ptr(), n64, k3d * norm2, p3z.data_ptr(), 0)), n64, 16 * n64)
