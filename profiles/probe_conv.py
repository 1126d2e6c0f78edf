"""Time the 2^26-sample ComplexF32 overlap-save convolution (4097 taps) with the library named by DSPB200_LIB
(timing probes: builds with -DDSP_PROBE=bits produce wrong results on purpose)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from dspb200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
n = 1 << 26
x = torch.view_as_complex(torch.randn(n, 2, device=dev))
y = torch.empty(n + bench.NV - 1, dtype=torch.complex64, device=dev)
plan = _lib.OsPlan(bench.make_taps(), 0)
fn = lambda: plan.exec_dev(x.data_ptr(), n, 1, y.data_ptr(), y.numel(), 0)   # noqa: E731
for _ in range(3):
    fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(st)
for _ in range(20):
    fn()
b.record(st)
torch.cuda.synchronize()
print(os.path.basename(os.environ.get("DSPB200_LIB", "libdspb200.so")), round(a.elapsed_time(b) / 20, 4), "ms")
