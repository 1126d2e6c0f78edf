"""Oracle: window functions (reference src/windows.jl:97-121 makewindow and callers). TEST INFRASTRUCTURE ONLY."""
import numpy as np


def _makewindow(winfunc, n):
    """src/windows.jl:97-121, non-zerophase, padding=0: winfunc on range(-0.5, 0.5; length=n); n==1 -> winfunc(0)."""
    if n < 0:
        raise ValueError("`n` must be nonnegative")
    if n == 1:
        return np.array([float(winfunc(np.float64(0.0)))])
    if n == 0:
        return np.zeros(0)
    # Julia range(-0.5, 0.5; length=n): value_i = -0.5 + (i-1)/(n-1), computed as a
    # twice-precision range whose end points are exact; linspace matches to <= 1 ulp.
    x = np.linspace(-0.5, 0.5, n)
    return np.asarray(winfunc(x), dtype=np.float64)


def rect(n):
    """src/windows.jl:142-144."""
    return _makewindow(lambda x: np.ones_like(x), n)


def hanning(n):
    """src/windows.jl:181-183: 0.5*(1+cospi(2x))."""
    return _makewindow(lambda x: 0.5 * (1 + _cospi(2 * x)), n)


def hamming(n):
    """src/windows.jl:206-208: muladd(0.46, cospi(2x), 0.54)."""
    return _makewindow(lambda x: 0.46 * _cospi(2 * x) + 0.54, n)


def bartlett(n):
    """src/windows.jl:380-382: 1 - abs(2x)."""
    return _makewindow(lambda x: 1 - np.abs(2 * x), n)


def kaiser(n, alpha):
    """src/windows.jl:600-605: I0(pi*alpha*sqrt(1-(2x)^2)) / I0(pi*alpha)."""
    pf = 1.0 / np.i0(np.pi * alpha)
    return _makewindow(lambda x: pf * np.i0(np.pi * alpha * np.sqrt(np.maximum(0.0, 1 - (2 * x) ** 2))), n)


def _cospi(t):
    """cospi(t) with exact zeros/ones at the quarter points like Julia's cospi."""
    t = np.asarray(t, dtype=np.float64)
    r = np.remainder(t, 2.0)
    out = np.cos(np.pi * r)
    out = np.where((r == 0.5) | (r == 1.5), 0.0, out)
    out = np.where(r == 0.0, 1.0, out)
    out = np.where(r == 1.0, -1.0, out)
    return out


def dpss(n, nw, ntapers=None):
    """dpss(n, nw, ntapers), src/windows.jl:668-720 (padding=0, zerophase=false): eigenvectors of the symmetric
    tridiagonal Slepian matrix for the `ntapers` largest eigenvalues; even-numbered tapers (2nd, 4th, ..) are signed
    so that their first nonzero sample is positive.  Returns an (n, ntapers) matrix (column = taper)."""
    import math
    from scipy.linalg import eigh_tridiagonal
    if ntapers is None:
        ntapers = math.ceil(2 * nw) - 1
    if not (0 < ntapers <= n):
        raise ValueError("ntapers must be in the interval (0, n]")
    if not (0 <= nw < n / 2):
        raise ValueError("nw must be in the interval [0, n/2)")
    v = float(_cospi(np.array([2 * nw / n]))[0])
    i = np.arange(n, dtype=np.float64)
    dv = v * ((n - 1) / 2 - i) ** 2
    j = np.arange(1, n, dtype=np.float64)
    ev = 0.5 * (j * n - j ** 2)
    _, vecs = eigh_tridiagonal(dv, ev, select="i", select_range=(n - ntapers, n - 1))
    rv = vecs[:, ::-1].copy()
    for c in range(1, ntapers, 2):          # 1-based even columns
        nz = np.flatnonzero(rv[:, c])
        if nz.size and rv[nz[0], c] < 0:
            rv[:, c] = -rv[:, c]
    return rv


def dpsseig(A, nw):
    """dpsseig(A, nw), src/windows.jl:739-775: concentration ratios of the tapers in the columns of A (output of dpss):
    q_i = 2w/nfft * sum_j seq[j] * autocorr_i[j], seq = [1, 2 sinc(2 w j)...], w = nw/n, autocorrelation by FFT of size
    nextfastfft(2n-1) with an unnormalised inverse (hence the /nfft)."""
    from .util import nextfastfft
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if not (0 <= nw < n / 2):
        raise ValueError("nw must be in the interval [0, n/2)")
    w = nw / n
    seq = np.empty(n)
    seq[0] = 1.0
    seq[1:] = 2 * np.sinc(2 * w * np.arange(1, n))
    nfft = nextfastfft(2 * n - 1)
    q = np.empty(A.shape[1])
    for i in range(A.shape[1]):
        tmp1 = np.zeros(nfft)
        tmp1[:n] = A[:, i]
        ac = np.fft.irfft(np.abs(np.fft.rfft(tmp1)) ** 2, nfft) * nfft      # brfft: unnormalised
        q[i] = 2 * w * float(np.dot(seq, ac[:n])) / nfft
    return q
