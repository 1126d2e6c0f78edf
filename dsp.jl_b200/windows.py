"""Window values needed host-side (reference src/windows.jl:97-121 makewindow; only the unpadded,
non-zerophase form that compute_window uses, src/periodograms.jl:248-257).  All return float64 vectors."""
import numpy as np

from .errors import ArgumentError


def _cospi(t):
    r = np.remainder(np.asarray(t, dtype=np.float64), 2.0)
    c = np.cos(np.pi * r)
    c[(r == 0.5) | (r == 1.5)] = 0.0
    c[r == 0.0] = 1.0
    c[r == 1.0] = -1.0
    return c


def makewindow(winfunc, n):
    """src/windows.jl:97-121 with padding=0, zerophase=false."""
    if n < 0:
        raise ArgumentError("`n` must be nonnegative")
    if n == 0:
        return np.zeros(0)
    if n == 1:
        return np.atleast_1d(np.asarray(winfunc(np.zeros(1)), dtype=np.float64))
    # Julia's range(-0.5, 0.5; length=n) is a twice-precision range: every point is the exact rational
    # -1/2 + i/(n-1) rounded once.  Extended precision reproduces that.
    xl = np.arange(n, dtype=np.longdouble) / np.longdouble(n - 1) - np.longdouble(0.5)
    x = xl.astype(np.float64)
    x[-1] = 0.5
    return np.asarray(winfunc(x), dtype=np.float64)


def rect(n):
    """src/windows.jl:142-144."""
    return makewindow(lambda x: np.ones_like(x), n)


def hanning(n):
    """src/windows.jl:181-183."""
    return makewindow(lambda x: 0.5 * (1 + _cospi(2 * x)), n)


hann = hanning


def hamming(n):
    """src/windows.jl:206-208."""
    return makewindow(lambda x: 0.46 * _cospi(2 * x) + 0.54, n)


def bartlett(n):
    """src/windows.jl:380-382."""
    return makewindow(lambda x: 1 - np.abs(2 * x), n)


def kaiser(n, alpha):
    """src/windows.jl:600-605."""
    pf = 1.0 / np.i0(np.pi * alpha)
    return makewindow(lambda x: pf * np.i0(np.pi * alpha * np.sqrt(np.clip(1 - (2 * x) ** 2, 0.0, None))), n)
