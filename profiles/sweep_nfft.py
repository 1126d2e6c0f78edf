"""Tuning sweep: every fused transform size, overlap-save (real / complex), Welch (real / complex) and STFT, 2^26 samples.
    DSPB200_LIB=<variant.so> python profiles/sweep_nfft.py [reps]      -> one JSON line per (kernel, nfft)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dspb200 import _lib  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda", 0)
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


n = 1 << 26
xr = torch.randn(n, device=dev)
xc = torch.view_as_complex(torch.randn(n, 2, device=dev))
yr = torch.empty(n, device=dev)
yc = torch.empty(n, dtype=torch.complex64, device=dev)
rng = np.random.default_rng(5)
for nfft in (256, 512, 1024, 2048, 4096, 8192, 16384):
    nb = nfft // 4 + 1
    br = rng.standard_normal(nb).astype(np.float32)
    bc = (rng.standard_normal(nb) + 1j * rng.standard_normal(nb)).astype(np.complex64)
    pr, pc = _lib.OsPlan(br, nfft), _lib.OsPlan(bc, nfft)
    assert pr.fused and pc.fused
    print(json.dumps({"kernel": "os_real", "nfft": nfft, "ms": round(timeit(lambda: pr.exec_dev(xr.data_ptr(), n, 1, yr.data_ptr(), n, 0)), 4)}))
    print(json.dumps({"kernel": "os_cplx", "nfft": nfft, "ms": round(timeit(lambda: pc.exec_dev(xc.data_ptr(), n, 1, yc.data_ptr(), n, 0)), 4)}))
    win = bench.hanning64(nfft)
    sr = _lib.SpecPlan(np.float32, nfft, nfft // 2, nfft, True, win)
    sc = _lib.SpecPlan(np.complex64, nfft, nfft // 2, nfft, False, win)
    pw = torch.empty(nfft, device=dev)
    k = sr.nsegments(n)
    print(json.dumps({"kernel": "welch_real", "nfft": nfft, "ms": round(timeit(lambda: sr.welch_dev(xr.data_ptr(), n, float(k), pw.data_ptr(), 0)), 4)}))
    print(json.dumps({"kernel": "welch_cplx", "nfft": nfft, "ms": round(timeit(lambda: sc.welch_dev(xc.data_ptr(), n, float(k), pw.data_ptr(), 0)), 4)}))
    # STFT (PSD) of 2^24 real samples, 50 % overlap
    m = 1 << 24
    ks = sr.nsegments(m)
    o = torch.empty((nfft // 2 + 1) * ks, device=dev)
    print(json.dumps({"kernel": "stft_real_psd", "nfft": nfft, "ms": round(timeit(lambda: sr.stft_dev(xr.data_ptr(), m, 1, 1.0, True, o.data_ptr(), 0)), 4)}))
    del o
