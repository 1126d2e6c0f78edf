"""Welch launch-configuration sweep (MODE x thread groups per CTA) at 2^26 samples, nfft = 4096 / 1024, 50 % overlap, hanning.
    python profiles/welch_cfg_sweep.py          (DSPB200_WELCH_CFG="mode,groups" is read by the library at every launch)"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dspb200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n = 1 << 26


def timeit(fn, reps=10):
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


for nfft in (4096, 1024):
    win = bench.hanning64(nfft)
    norm2 = float(np.sum(win * win))
    for name, dt, x in (("F32", np.float32, torch.randn(n, device=dev)),
                        ("CF32", np.complex64, torch.view_as_complex(torch.randn(n, 2, device=dev)))):
        plan = _lib.SpecPlan(dt, nfft, nfft // 2, nfft, dt == np.float32, win)
        p = torch.empty(nfft, dtype=torch.float32, device=dev)
        r = plan.nsegments(n) * norm2
        ref = None
        for cfg in ("", "2,1", "2,2", "2,3", "3,1", "3,2", "1,1", "1,2", "1,3"):
            if cfg:
                os.environ["DSPB200_WELCH_CFG"] = cfg
            else:
                os.environ.pop("DSPB200_WELCH_CFG", None)
            try:
                ms = timeit(lambda: plan.welch_dev(x.data_ptr(), n, r, p.data_ptr(), 0))
                out = p.cpu().numpy().copy()
                if ref is None:
                    ref = out
                err = float(np.linalg.norm(out - ref) / np.linalg.norm(ref))
                print(json.dumps({"nfft": nfft, "dtype": name, "cfg": cfg or "auto", "ms": round(ms, 4), "relerr_vs_auto": err}))
            except Exception as e:      # configuration does not fit
                print(json.dumps({"nfft": nfft, "dtype": name, "cfg": cfg, "error": str(e)[:80]}))
        os.environ.pop("DSPB200_WELCH_CFG", None)
