"""Multitaper spectral estimation front ends (reference src/multitaper.jl:5-404), backed by
libdspb200: `dpss`, `MTConfig`, `mt_pgram`, `mt_spectrogram`."""
import math

import numpy as np

from . import _lib
from .errors import ArgumentError, DimensionMismatch, DomainError
from .periodograms import Periodogram, Spectrogram, arraysplit_count
from .util import fftabs2type, fftfreq, fftintype, nextfastfft, rfftfreq


def dpss(n, nw, ntapers=None):
    """dpss(n, nw, ntapers=ceil(2nw)-1), src/windows.jl:668-720: Slepian tapers as an (n, ntapers) matrix with unit-norm
    columns (symmetric tapers with positive mean, antisymmetric ones starting positive)."""
    from scipy.signal.windows import dpss as _dpss
    ntapers = math.ceil(2 * nw) - 1 if ntapers is None else int(ntapers)
    if not (0 < ntapers <= n):
        raise DomainError("ntapers must be in the interval (0, n]")
    if not (0 <= nw < n / 2):
        raise DomainError("nw must be in the interval [0, n/2)")
    w = _dpss(int(n), nw, ntapers, sym=True, norm=2)
    return np.ascontiguousarray(np.atleast_2d(w).T)


class MTConfig:
    """MTConfig{T}(n_samples; fs, nfft, window, nw, ntapers, taper_weights, onesided), src/multitaper.jl:117-141."""

    def __init__(self, eltype, n_samples, fs=1, nfft=None, window=None, nw=4, ntapers=None, taper_weights=None,
                 onesided=None, noverlap=0):
        eltype = np.dtype(eltype)
        cplx = eltype.kind == "c"
        onesided = (not cplx) if onesided is None else bool(onesided)
        if onesided and cplx:
            raise ArgumentError("cannot compute one-sided FFT of a complex signal")
        if n_samples <= 0:
            raise ArgumentError("`n_samples` must be positive")
        nfft = (1 << (int(n_samples) - 1).bit_length()) if nfft is None else int(nfft)     # nextpow(2, n_samples)
        if nfft < n_samples:
            raise ArgumentError("Must have `nfft >= n_samples`")
        ntapers = math.ceil(2 * nw) - 1 if ntapers is None else int(ntapers)
        if window is None:
            win = dpss(n_samples, nw, ntapers)
            norm2 = np.ones(ntapers)
        else:
            win = np.asarray(window, dtype=np.float64)
            if win.ndim != 2 or win.shape[0] != n_samples:
                raise DimensionMismatch("window must be an n_samples x ntapers matrix")
            ntapers = win.shape[1]
            norm2 = np.sum(win * win, axis=0)
        w = np.full(ntapers, 1.0 / ntapers) if taper_weights is None else np.asarray(taper_weights, dtype=np.float64)
        self.n_samples, self.nfft, self.ntapers, self.fs, self.onesided = int(n_samples), nfft, ntapers, fs, onesided
        self.window = win
        self.r = fs * norm2 / w                                                      # :135-139
        self.freq = rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs)
        self.intype = fftintype(eltype)
        scaled = (win / np.sqrt(self.r)[None, :]).T                                  # rows pre-scaled by 1/sqrt(r_t)
        self.plan = _lib.MtPlan(self.intype, n_samples, noverlap, nfft, onesided, scaled)


def mt_pgram(s, config=None, onesided=None, nfft=None, fs=1, nw=4, ntapers=None, window=None):
    """mt_pgram(s; onesided, nfft=nextfastfft(length(s)), fs, nw, ntapers, window) / mt_pgram(s, config),
    src/multitaper.jl:178-242."""
    s = np.asarray(s)
    if s.ndim != 1:
        raise ArgumentError("expected a vector")
    if config is None:
        config = MTConfig(s.dtype, s.size, fs=fs, nfft=nextfastfft(s.size) if nfft is None else nfft, window=window, nw=nw,
                          ntapers=ntapers, onesided=onesided)
    if s.size != config.n_samples:
        raise DimensionMismatch("Expected `signal` to be of length `config.n_samples`")
    sig = np.ascontiguousarray(s, dtype=config.intype)
    out = np.empty(config.plan.nout, dtype=fftabs2type(config.intype))
    config.plan.mt_pgram(sig, out)
    return Periodogram(out, config.freq)


def mt_spectrogram(s, n=None, n_overlap=None, fs=1, onesided=None, nfft=None, nw=4, ntapers=None, window=None):
    """mt_spectrogram(signal, n, n_overlap; fs, onesided, kwargs...), src/multitaper.jl:262-404 (default nfft = nextpow(2, n))."""
    s = np.asarray(s)
    if s.ndim != 1:
        raise ArgumentError("expected a vector")
    n = s.size >> 3 if n is None else int(n)
    n_overlap = n >> 1 if n_overlap is None else int(n_overlap)
    if n <= n_overlap:
        raise ArgumentError("Need `samples_per_window > n_overlap_samples`")
    config = MTConfig(s.dtype, n, fs=fs, nfft=nfft, window=window, nw=nw, ntapers=ntapers, onesided=onesided, noverlap=n_overlap)
    k = arraysplit_count(s.size, n, n_overlap)
    sig = np.ascontiguousarray(s, dtype=config.intype)
    out = np.zeros((config.plan.nout, k), dtype=fftabs2type(config.intype), order="F")
    if k > 0:
        config.plan.mt_spectrogram(sig, out)
    t = (n / 2 + (n - n_overlap) * np.arange(k, dtype=np.float64)) / fs
    return Spectrogram(out, config.freq, t)
