"""Oracle: FFT type maps and nextfastfft (reference src/util.jl:92-135). TEST INFRASTRUCTURE ONLY."""
import numpy as np

_FFTW_REAL = (np.float32, np.float64)
_FFTW_CPLX = (np.complex64, np.complex128)


def fftintype(dt):
    """src/util.jl:92-94 -- FFTW number types pass through, other reals -> Float64, complex -> ComplexF64."""
    dt = np.dtype(dt).type
    if dt in _FFTW_REAL or dt in _FFTW_CPLX:
        return np.dtype(dt)
    return np.dtype(np.complex128) if np.issubdtype(dt, np.complexfloating) else np.dtype(np.float64)


def fftouttype(dt):
    """src/util.jl:97-99."""
    dt = np.dtype(dt).type
    if dt in _FFTW_CPLX:
        return np.dtype(dt)
    if dt in _FFTW_REAL:
        return np.dtype(np.complex64 if dt is np.float32 else np.complex128)
    return np.dtype(np.complex128)


def fftabs2type(dt):
    """src/util.jl:102-104."""
    dt = np.dtype(dt).type
    if dt in (np.float32, np.complex64):
        return np.dtype(np.float32)
    return np.dtype(np.float64)


def nextfastfft(n):
    """src/util.jl:107,134 -- nextprod((2,3,5,7), n): smallest 2^a 3^b 5^c 7^d >= n."""
    n = int(n)
    if n <= 1:
        return 1
    best = None
    p7 = 1
    while p7 < 2 * n:
        p5 = p7
        while p5 < 2 * n:
            p3 = p5
            while p3 < 2 * n:
                p2 = p3
                while p2 < n:
                    p2 *= 2
                if best is None or p2 < best:
                    best = p2
                p3 *= 3
            p5 *= 5
        p7 *= 7
    return best


def rfftfreq(n, fs=1.0):
    """FFTW.rfftfreq(n, fs) = (0:n>>1) * fs/n (used at src/periodograms.jl:573,834)."""
    return np.arange(n // 2 + 1) * (fs / n)


def fftfreq(n, fs=1.0):
    """FFTW.fftfreq(n, fs): [0..ceil(n/2)-1, -floor(n/2)..-1] * fs/n."""
    k = np.arange(n)
    k = np.where(k < (n + 1) // 2, k, k - n)
    return k * (fs / n)


def hilbert(x):
    """src/util.jl:31-75 -- analytic signal along dim 1: X = zeros(n); X[1:n>>1+1] = rfft(x); X[2:n÷2+isodd(n)] *= 2;
    ifft(X).  Float32 in -> ComplexF32 out, every other real eltype through Float64 (:43)."""
    x = np.asarray(x)
    tin = fftintype(x.dtype)
    tout = fftouttype(tin)
    a = x.astype(tin, copy=False)
    n = a.shape[0]
    X = np.zeros(a.shape, dtype=np.complex128)
    X[: (n >> 1) + 1] = np.fft.rfft(a.astype(np.float64), axis=0)
    X[1: n // 2 + (n & 1)] *= 2.0
    return np.fft.ifft(X, axis=0).astype(tout)
