"""Host-side scalar helpers mirroring reference src/util.jl:92-135 (FFT type maps, nextfastfft)."""
import numpy as np

_REAL = (np.dtype(np.float32), np.dtype(np.float64))
_CPLX = (np.dtype(np.complex64), np.dtype(np.complex128))


def fftintype(dt):
    """src/util.jl:92-94."""
    dt = np.dtype(dt)
    if dt in _REAL or dt in _CPLX:
        return dt
    return np.dtype(np.complex128) if dt.kind == "c" else np.dtype(np.float64)


def fftouttype(dt):
    """src/util.jl:97-99."""
    dt = np.dtype(dt)
    if dt in _CPLX:
        return dt
    if dt == np.float32:
        return np.dtype(np.complex64)
    return np.dtype(np.complex128)


def fftabs2type(dt):
    """src/util.jl:102-104."""
    dt = np.dtype(dt)
    return np.dtype(np.float32) if dt in (np.dtype(np.float32), np.dtype(np.complex64)) else np.dtype(np.float64)


def nextfastfft(n):
    """src/util.jl:134 -- nextprod((2,3,5,7), n).  Tuples map element-wise (:135)."""
    if isinstance(n, (tuple, list)):
        return tuple(nextfastfft(v) for v in n)
    n = int(n)
    if n <= 1:
        return 1
    best = 1 << (n - 1).bit_length()
    f7 = 1
    while f7 < best:
        f5 = f7
        while f5 < best:
            f3 = f5
            while f3 < best:
                v = f3
                while v < n:
                    v *= 2
                best = min(best, v)
                f3 *= 3
            f5 *= 5
        f7 *= 7
    return best


def rfftfreq(n, fs=1.0):
    """FFTW.rfftfreq: (0 : n>>1) * fs / n (src/periodograms.jl:573, 834)."""
    return np.arange(n // 2 + 1, dtype=np.float64) * (fs / n)


def fftfreq(n, fs=1.0):
    """FFTW.fftfreq: non-negative bins first, then the negative ones."""
    k = np.arange(n, dtype=np.int64)
    k[(n + 1) // 2:] -= n
    return k * (fs / n)
