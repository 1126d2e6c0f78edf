"""Thin clients of the GPU convolution / FIR / FFT path: xcorr, FIR filtfilt, finddelay,
shiftsignal, alignsignals, hilbert.  Each is a few host lines over `conv` / `filt_` / one FFT pair exactly as in the
reference."""
import numpy as np

from . import _lib
from .device import DeviceArray
from .dspbase import _cols, _promote, conv
from .util import fftintype, fftouttype
from .errors import ArgumentError, DimensionMismatch, DomainError
from .filters import filt_ as _filt_hx_


def xcorr(u, v=None, padmode="none", scaling="none"):
    """xcorr(u[, v]; padmode, scaling), src/dspbase.jl:867-898: conv(u, reverse(conj(v))) -- conjugates the SECOND
    argument (MATLAB / scipy convention)."""
    u = np.asarray(u)
    v = u if v is None else np.asarray(v)
    if u.ndim != 1 or v.ndim != 1:
        raise ArgumentError("xcorr takes vectors")
    su, sv = u.size, v.size
    padmode = padmode.lstrip(":") if isinstance(padmode, str) else padmode
    scaling = scaling.lstrip(":") if isinstance(scaling, str) else scaling
    if scaling == "biased" and su != sv:
        raise DimensionMismatch("scaling only valid for vectors of same length")
    if padmode == "longest":
        if su < sv:
            u = np.concatenate([u, np.zeros(sv - su, dtype=u.dtype)])
        elif sv < su:
            v = np.concatenate([v, np.zeros(su - sv, dtype=v.dtype)])
    elif padmode != "none":
        raise ArgumentError("padmode keyword argument must be either :none or :longest")
    res = conv(u, np.conj(v)[::-1])
    if scaling == "biased":
        res = res / su
    elif scaling != "none":
        raise ArgumentError("scaling keyword argument must be either :none or :biased")
    return res


def _extrapolate_signal(sig, pad_length):
    """extrapolate_signal!, src/Filters/filt.jl:245-259: odd-symmetric extension of both ends."""
    n = sig.shape[0]
    # explicit indices: a slice stop of -1 (pad_length == n - 1, i.e. len(x) == len(b)) would mean "last element"
    head = 2 * sig[0] - sig[np.arange(pad_length, 0, -1)]
    tail = 2 * sig[n - 1] - sig[np.arange(n - 2, n - 2 - pad_length, -1)]
    return np.concatenate([head, sig, tail], axis=0)


def filtfilt(b, a_or_x, x=None):
    """filtfilt(b, x) / filtfilt(b, a, x) with length(a) == 1, src/Filters/filt.jl:301-337: zero-phase FIR filtering --
    the signal is extended odd-symmetrically by nb-1 samples and filtered once with conv(b, reverse(b))."""
    b = np.asarray(b)
    if x is None:
        x = np.asarray(a_or_x)
    else:
        a = np.atleast_1d(np.asarray(a_or_x))
        x = np.asarray(x)
        if a.size != 1:
            raise NotImplementedError("IIR filtfilt is outside the B200 hot-path scope (serial recurrence)")
        if a[0] != 1:
            b = b / a[0]
    nb = b.size
    if nb == 0:
        raise ArgumentError("filter vector b must be non-empty")
    if x.shape[0] < nb:
        raise ArgumentError("signal must be at least as long as the filter")     # BoundsError in the reference
    T = _promote(b, x)
    if T.kind in "biu":
        T = np.dtype(np.float64)
    bT = b.astype(T)
    newb = np.convolve(bT, bT[::-1])             # filt!(newb, b, reverse(b)) mirrored, :309-314 (2nb-1 taps, tiny, host)
    ext = _extrapolate_signal(x.astype(T), nb - 1)
    out = np.empty(ext.shape, dtype=T, order="F")
    _filt_hx_(out, newb.astype(T), ext)          # filt!(extrapolated, newb, extrapolated), :322
    return out[2 * nb - 2:]                       # drop garbage at start, :325


def finddelay(x, y):
    """finddelay(x, y), src/util.jl:360-368."""
    x = np.asarray(x)
    y = np.asarray(y)
    s = xcorr(y, x, padmode="none")
    mag = np.abs(s)
    idxs = np.flatnonzero(mag == mag.max()) + 1          # 1-based like the reference
    center = x.size
    return int(center - idxs[np.argmin(np.abs(center - idxs))])


def shiftsignal(x, s):
    """shiftsignal(x, s), src/util.jl:379-412."""
    x = np.array(x, copy=True)
    n = x.size
    if abs(s) > n:
        raise DomainError("The absolute value of s must not be greater than the length of x")
    if s > 0:
        x[s:] = x[:n - s].copy()
        x[:s] = 0
    elif s < 0:
        x[:n + s] = x[-s:].copy()
        x[n + s:] = 0
    return x


def alignsignals(x, y):
    """alignsignals(x, y), src/util.jl:419-427."""
    d = finddelay(x, y)
    return shiftsignal(x, -d), d


def hilbert(x):
    """hilbert(x), src/util.jl:31-75: analytic signal x + j*H{x} of a real signal along the first dimension (every
    trailing index is an independent column).  Float32 stays single precision, every other real type is computed in
    Float64 (fftintype, src/util.jl:43, 92-94).  A `DeviceArray` is transformed in HBM and a `DeviceArray` returned."""
    if isinstance(x, DeviceArray):
        if x.dtype.kind != "f":
            raise ArgumentError("hilbert takes a real signal")
        n = x.shape[0] if x.ndim else 1
        ncols = x.size // max(n, 1)
        out = DeviceArray(x.shape, fftouttype(x.dtype))
        if x.size:
            _lib.hilbert_dev(x.dtype, x.ptr, n, ncols, out.ptr, 0)
        return out
    x = np.asarray(x)
    if x.dtype.kind == "c":
        raise ArgumentError("hilbert takes a real signal")
    tin = fftintype(x.dtype)
    a, n, ncols = _cols(x, tin)
    res = np.empty((n, ncols), dtype=fftouttype(tin), order="F")
    if a.size:
        _lib.hilbert(a, n, ncols, res)
    return res.reshape(x.shape)      # column c of res <-> trailing index c (C order over the trailing dims)
