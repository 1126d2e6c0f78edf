"""Wall-clock timing of the widened host-pointer entry points: first call (plan creation + allocation) and the
mean of the following calls (cached cuFFT plans and scratch).    python profiles/time_widened.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import dspb200 as dsp  # noqa: E402

rng = np.random.default_rng(5)


def run(name, fn, reps=5):
    t0 = time.perf_counter()
    fn()
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    rest = (time.perf_counter() - t0) / reps
    print(json.dumps({"op": name, "first_call_ms": round(first * 1e3, 3), "next_calls_ms": round(rest * 1e3, 3)}))


u = rng.standard_normal(1 << 20).astype(np.float32)
v = rng.standard_normal(1 << 12).astype(np.float32)
run("conv(u[2^20], v[2^12]; algorithm=:fft_simple) F32", lambda: dsp.conv(u, v, algorithm="fft_simple"))
a = rng.standard_normal((1024, 1024)).astype(np.float32)
b = rng.standard_normal((63, 63)).astype(np.float32)
run("conv(A[1024x1024], B[63x63]) F32", lambda: dsp.conv(a, b))
run("conv(A[1024x1024], B[63x63]; algorithm=:fft_simple) F32", lambda: dsp.conv(a, b, algorithm="fft_simple"))
run("conv(A[1024x1024], B[63x63]; algorithm=:fft_overlapsave) F32 (batched blocks)", lambda: dsp.conv(a, b, algorithm="fft_overlapsave"))
ad, bd = dsp.to_device(a), dsp.to_device(b)
run("conv(A, B; algorithm=:fft_overlapsave) F32, device-resident", lambda: dsp.conv(ad, bd, algorithm="fft_overlapsave"))
run("conv(A, B; algorithm=:fft_simple) F32, device-resident", lambda: dsp.conv(ad, bd, algorithm="fft_simple"))
big = rng.standard_normal((8192, 8192)).astype(np.float32)
run("conv(A[8192x8192], B[63x63]; algorithm=:fft_overlapsave) F32", lambda: dsp.conv(big, b, algorithm="fft_overlapsave"), reps=2)
c3 = rng.standard_normal((128, 128, 64))
d3 = rng.standard_normal((9, 9, 9))
run("conv(A[128x128x64], B[9x9x9]) F64", lambda: dsp.conv(c3, d3))
run("conv(A[128x128x64], B[9x9x9]; algorithm=:fft_simple) F64", lambda: dsp.conv(c3, d3, algorithm="fft_simple"))
run("conv(A[128x128x64], B[9x9x9]; algorithm=:fft_overlapsave) F64", lambda: dsp.conv(c3, d3, algorithm="fft_overlapsave"))
m = rng.standard_normal((2048, 2048)).astype(np.float32)
run("periodogram(s[2048x2048]) F32", lambda: dsp.periodogram(m))
run("periodogram(s[2048x2048]; radialavg=true) F32", lambda: dsp.periodogram(m, radialavg=True))
md = dsp.to_device(m)
run("periodogram(s[2048x2048]) F32, device-resident", lambda: dsp.periodogram(md))
x = rng.standard_normal(1 << 22)
run("hilbert(x[2^22]) F64", lambda: dsp.hilbert(x))
run("xcorr(u[2^20], v[2^12]) F32", lambda: dsp.xcorr(u, v))
s = rng.standard_normal(1 << 16).astype(np.float32)
run("mt_pgram(s[2^16]; nw=4) F32", lambda: dsp.mt_pgram(s))
s2 = rng.standard_normal(1 << 22).astype(np.float32)
run("mt_spectrogram(s[2^22], 1024, 512; nw=4) F32", lambda: dsp.mt_spectrogram(s2, 1024, 512))
sig = rng.standard_normal((8, 1 << 14))
run("mt_coherence(signal[8 x 2^14]) F64", lambda: dsp.mt_coherence(sig))
cfg = dsp.MTCrossSpectraConfig(8, 1 << 14, eltype=np.float64)
sigd = dsp.to_device(sig)
run("mt_coherence(signal[8 x 2^14], config) F64, device-resident, config reused", lambda: dsp.mt_coherence(sigd, cfg))
