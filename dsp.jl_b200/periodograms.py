"""periodogram / welch_pgram / spectrogram / stft front ends with the reference's signatures
(src/periodograms.jl), backed by libdspb200.  `welch_pgram!` is spelled `welch_pgram_`."""
import warnings

import numpy as np

from . import _lib
from .device import DeviceArray
from .errors import ArgumentError, DimensionMismatch, DomainError
from .util import fftabs2type, fftfreq, fftintype, fftouttype, nextfastfft, rfftfreq

_UNSET = object()


# --------------------------------------------------------------------------------------------- result types

class Periodogram:
    """Periodogram(power, freq), src/periodograms.jl:270-273."""

    def __init__(self, power, freq):
        self.power = power
        self.freq = freq


class Periodogram2:
    """Periodogram2(power, freq1, freq2), src/periodograms.jl:300-304: two-dimensional PSD; `freq` is the pair."""

    def __init__(self, power, freq1, freq2):
        self.power, self.freq1, self.freq2 = power, freq1, freq2

    @property
    def freq(self):
        return (self.freq1, self.freq2)


class Spectrogram:
    """Spectrogram(power, freq, time), src/periodograms.jl:773-777."""

    def __init__(self, power, freq, time):
        self.power = power
        self.freq = freq
        self.time = time


def power(p):
    """src/periodograms.jl:310."""
    return p.power


def freq(p):
    """src/periodograms.jl:329-333."""
    return p.freq


def time(p):
    """src/periodograms.jl:793."""
    return p.time


# --------------------------------------------------------------------------------------------- window / segmenting

def compute_window(window, n):
    """src/periodograms.jl:248-257 -> (window values as float64 or None, norm2)."""
    if window is None:
        return None, float(n)
    if callable(window):
        win = np.asarray(window(n), dtype=np.float64)
        return win, float(np.sum(win * win))
    win = np.asarray(window)
    if win.ndim != 1 or win.size != n:
        raise DimensionMismatch("length of window must match input")
    if np.iscomplexobj(win):
        raise NotImplementedError("complex windows are outside the B200 hot-path scope")
    return np.asarray(win, dtype=np.float64), float(np.sum(np.abs(win) ** 2))


def arraysplit_count(length, n, noverlap):
    """ArraySplit.k, src/periodograms.jl:49-50."""
    if not (0 <= noverlap < n):
        raise DomainError("noverlap must be between zero and n")     # :44
    return (length - n) // (n - noverlap) + 1 if length >= n else 0


def arraysplit(s, n, noverlap, nfft=None, window=None):
    """arraysplit(s, n, noverlap, nfft=n, window=nothing), src/periodograms.jl:134-137 (ArraySplit :32-73).
    Returns all k segments at once as a (k, nfft) array (`collect` of the reference's lazy iterator, :137):
    row i = [window .* s[i*hop : i*hop+n], zeros(nfft - n)], eltype fftintype(eltype(s))."""
    s = np.asarray(s)
    nfft = n if nfft is None else int(nfft)
    if not (0 <= noverlap < n):
        raise DomainError("noverlap must be between zero and n")     # :44
    if nfft < n:
        raise DomainError("nfft must be >= n")                      # :45
    sig = _signal(s)
    k = arraysplit_count(sig.size, n, noverlap)
    out = np.zeros((k, nfft), dtype=sig.dtype)
    if k == 0:
        return out
    win = None
    if window is not None:
        win = np.asarray(window)
        if win.size != n:
            raise DimensionMismatch("length of window must match input")
        win = np.asarray(win, dtype=np.float64)
    plan = _lib.SpecPlan(sig.dtype, n, noverlap, nfft, sig.dtype.kind != "c", win)
    plan.arraysplit(sig, out)
    plan.close()
    return out


def fftshift(p):
    """FFTW.fftshift(::Periodogram / ::Spectrogram), src/periodograms.jl:331-333, 778-780: two-sided spectra are
    rotated so frequencies ascend; one-sided ones are returned unchanged."""
    if isinstance(p, Periodogram2):                                                 # :336-337
        return Periodogram2(np.fft.fftshift(p.power), np.fft.fftshift(p.freq1), np.fft.fftshift(p.freq2))
    f = np.asarray(p.freq)
    if f.size == 0 or np.all(np.diff(f) > 0):
        return p
    if isinstance(p, Spectrogram):
        return Spectrogram(np.fft.fftshift(p.power, axes=0), np.fft.fftshift(f), p.time)
    return Periodogram(np.fft.fftshift(p.power), np.fft.fftshift(f))


def _signal(s):
    s = np.asarray(s)
    if s.ndim != 1:
        raise ArgumentError("expected a vector (the 2-D periodogram is outside the B200 hot-path scope)")
    return np.ascontiguousarray(s, dtype=fftintype(s.dtype))           # buffer eltype, :55


# --------------------------------------------------------------------------------------------- WelchConfig

class WelchConfig:
    """WelchConfig(data_or_nsamples, eltype; n, noverlap, onesided, nfft, fs, window), src/periodograms.jl:516-587.
    Owns the device plan (segmenter + FFT + window), reusable across calls like the reference's plan/buffers."""

    def __init__(self, data, eltype=None, n=None, noverlap=None, onesided=None, nfft=None, fs=1, window=_UNSET):
        if eltype is None:
            data = np.asarray(data)
            nsamples, eltype = data.shape[-1], data.dtype
        else:
            nsamples = int(data)
        eltype = np.dtype(eltype)
        n = nsamples >> 3 if n is None else int(n)
        noverlap = n >> 1 if noverlap is None else int(noverlap)
        cplx = eltype.kind == "c"
        onesided = (not cplx) if onesided is None else bool(onesided)
        nfft = nextfastfft(n) if nfft is None else int(nfft)
        if window is _UNSET:                                           # :582-587
            warnings.warn("Omitting `window` is deprecated; specify `window=None` for the old behaviour or "
                          "`window=hanning` for the future default.", DeprecationWarning, stacklevel=3)
            window = None
        if onesided and cplx:
            raise ArgumentError("cannot compute one-sided FFT of a complex signal")   # :564
        if nfft < n:
            raise DomainError("nfft must be >= n")                     # :565
        win, norm2 = compute_window(window, n)
        if not (0 <= noverlap < n):
            raise DomainError("noverlap must be between zero and n")   # ArraySplit :44
        self.nsamples, self.noverlap, self.onesided, self.nfft, self.fs = n, noverlap, onesided, nfft, fs
        self.window = win
        self.r = fs * norm2                                            # :568
        self.intype = fftintype(eltype)                                # eltype(inbuf) = float(T), :569
        self.freq = rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs)   # :573
        self.plan = _lib.SpecPlan(self.intype, n, noverlap, nfft, onesided, win)


def welch_pgram(s, n=None, noverlap=None, onesided=None, nfft=None, fs=1, window=_UNSET):
    """welch_pgram(s, n, noverlap; kw...) (src/periodograms.jl:647-649) or welch_pgram(s, config) (:702-705)."""
    if isinstance(n, WelchConfig):
        config = n
    else:
        if isinstance(s, DeviceArray):
            nn = s.shape[-1] >> 3 if n is None else int(n)
            config = WelchConfig(s.shape[-1], s.dtype, n=nn, noverlap=(nn >> 1 if noverlap is None else noverlap),
                                 onesided=onesided, nfft=nfft, fs=fs, window=window)
        else:
            s = np.asarray(s)
            nn = s.shape[-1] >> 3 if n is None else int(n)
            config = WelchConfig(s, n=nn, noverlap=(nn >> 1 if noverlap is None else noverlap), onesided=onesided,
                                 nfft=nfft, fs=fs, window=window)
    if isinstance(s, DeviceArray):
        return _welch_device(s, config)
    sig = _signal(s)
    out = np.empty(config.nfft // 2 + 1 if config.onesided else config.nfft, dtype=fftabs2type(sig.dtype))
    return _welch_helper(out, sig, config)


def filt_welch(x, n_or_b, b_or_config=None, config=None, nfft=None):
    """welch_pgram(filt(b, x), config) as one pipelined call: filt_welch(x, b, config) for a host array, or
    filt_welch(ptr, n, b, config) for a raw (ideally pinned) host pointer to `n` samples of eltype config.intype.
    The composition of the reference's `filt(b, x)` (src/dspbase.jl:14-15, FFT path src/Filters/filt.jl:445-521) and
    `welch_pgram(y, config)` (src/periodograms.jl:702-759); the stream is uploaded in chunks that overlap the kernels and
    the filter output never leaves the GPU (dspb200_filt_welch_exec).  Returns the Periodogram of the filtered stream."""
    if isinstance(x, (int, np.integer)):
        ptr_, n, b, cfg = int(x), int(n_or_b), b_or_config, config
        keep = None
    else:
        b, cfg = n_or_b, b_or_config
        keep = _signal(x)
        if keep.dtype != cfg.intype:
            raise ArgumentError(f"float(eltype(s)) = {keep.dtype} doesn't match the eltype of the input buffer: {cfg.intype}.")
        ptr_, n = _lib.ptr(keep), keep.size
    if not isinstance(cfg, WelchConfig):
        raise ArgumentError("filt_welch needs a WelchConfig")
    taps = np.ascontiguousarray(np.asarray(b), dtype=cfg.intype)
    if taps.ndim != 1 or taps.size == 0:
        raise ArgumentError("filter vector b must be non-empty")
    out = np.zeros(cfg.nfft // 2 + 1 if cfg.onesided else cfg.nfft, dtype=fftabs2type(cfg.intype))
    k = arraysplit_count(n, cfg.nsamples, cfg.noverlap)
    if k > 0:
        key = (taps.tobytes(), nfft)
        plans = cfg.__dict__.setdefault("_os_plans", {})
        osp = plans.get(key)
        if osp is None:
            osp = plans[key] = _lib.OsPlan(taps, 0 if nfft is None else int(nfft))
        cfg.plan.filt_welch_ptr(osp, ptr_, n, k * cfg.r, _lib.ptr(out))
    return Periodogram(out, cfg.freq)


def welch_pgram_(out, s, n=None, noverlap=None, onesided=None, nfft=None, fs=1, window=None):
    """welch_pgram!(out, s, config) (src/periodograms.jl:734-744) / welch_pgram!(out, s, n, noverlap; kw...) (:683-686)."""
    if isinstance(n, WelchConfig):
        config = n
    else:
        s0 = np.asarray(s)
        nn = s0.shape[-1] >> 3 if n is None else int(n)
        config = WelchConfig(s0, n=nn, noverlap=(nn >> 1 if noverlap is None else noverlap), onesided=onesided,
                             nfft=nfft, fs=fs, window=window)
    sdt = np.asarray(s).dtype
    if out.size != config.freq.size:
        raise DimensionMismatch(f"Expected `output` to be of length `length(config.freq)`; got {out.size} and {config.freq.size}")
    if out.dtype != fftabs2type(sdt):
        raise ArgumentError(f"Eltype of output ({out.dtype}) doesn't match the expected type: {fftabs2type(sdt)}.")
    if fftintype(sdt) != config.intype:
        raise ArgumentError(f"float(eltype(s)) = {sdt} doesn't match the eltype of the input buffer: {config.intype}.")
    return _welch_helper(out, _signal(s), config)


def _welch_helper(out, sig, config):
    """welch_pgram_helper!, src/periodograms.jl:746-759 (the segment loop runs on the GPU)."""
    if sig.dtype != config.intype:
        raise ArgumentError(f"float(eltype(s)) = {sig.dtype} doesn't match the eltype of the input buffer: {config.intype}.")
    k = arraysplit_count(sig.size, config.nsamples, config.noverlap)
    if k == 0:
        out[...] = 0
        return Periodogram(out, config.freq)
    r = k * config.r                                                   # :751
    res = out if out.flags.c_contiguous else np.empty(out.shape, dtype=out.dtype)
    config.plan.welch(sig, r, res)
    if res is not out:
        out[...] = res
    return Periodogram(out, config.freq)


def _welch_device(s, config):
    """welch_pgram on a device-resident vector: only the nout-sample power vector crosses PCIe."""
    if s.ndim != 1:
        raise ArgumentError("expected a vector")
    if s.dtype != config.intype:
        raise ArgumentError(f"float(eltype(s)) = {s.dtype} doesn't match the eltype of the input buffer: {config.intype}.")
    out = np.zeros(config.nfft // 2 + 1 if config.onesided else config.nfft, dtype=fftabs2type(s.dtype))
    k = arraysplit_count(s.shape[0], config.nsamples, config.noverlap)
    if k == 0:
        return Periodogram(out, config.freq)
    dout = DeviceArray(out.shape, out.dtype)
    config.plan.welch_dev(s.ptr, s.shape[0], k * config.r, dout.ptr, 0)
    dout.to_host(out)
    return Periodogram(out, config.freq)


def _periodogram2(s, nfft, fs, radialsum, radialavg):
    """periodogram(s::AbstractMatrix; nfft, fs, radialsum, radialavg), src/periodograms.jl:473-509."""
    if s.dtype.kind == "c":
        raise ArgumentError("the periodogram of a matrix takes a real signal")
    nfft = tuple(nextfastfft(n) for n in s.shape) if nfft is None else tuple(int(n) for n in nfft)
    if not (s.shape[0] <= nfft[0] and s.shape[1] <= nfft[1]):
        raise ArgumentError("nfft must be >= size(s)")
    if not (s.shape[0] > 1 and s.shape[1] > 1):
        raise ArgumentError("dimensions of s must be > 1")
    if radialsum and radialavg:
        raise ArgumentError("radialsum and radialavg are mutually exclusive")
    r = fs * s.size                                                                # fs * norm2, :491
    nmin = min(nfft)
    radial_freq = lambda n: np.arange(n, dtype=np.float64) * (fs / nmin)           # Frequencies(n, n, fs / nmin), :507
    ptype = 1 if radialsum else 2 if radialavg else 0
    if isinstance(s, DeviceArray):                                                 # device-resident matrix: result stays in HBM
        if s.dtype.kind != "f":
            raise ArgumentError("device matrices must be Float32 or Float64")
        T = fftabs2type(s.dtype)
        dout = DeviceArray(nfft if ptype == 0 else ((nmin >> 1) + 1,), T)
        _lib.periodogram2_dev(s.dtype, s.ptr, s.shape, nfft, r, ptype, dout.ptr)
        if ptype == 0:
            return Periodogram2(dout, fftfreq(nfft[0], fs), fftfreq(nfft[1], fs))
        return Periodogram(dout, radial_freq(dout.shape[0]))
    sig = np.asfortranarray(s, dtype=fftintype(s.dtype))
    T = fftabs2type(sig.dtype)
    if ptype == 0:
        out = np.empty(nfft, dtype=T, order="F")
        _lib.periodogram2(sig, nfft, r, 0, out)
        return Periodogram2(out, fftfreq(nfft[0], fs), fftfreq(nfft[1], fs))
    out = np.empty((nmin >> 1) + 1, dtype=T)
    _lib.periodogram2(sig, nfft, r, ptype, out)
    return Periodogram(out, radial_freq(out.size))


def periodogram(s, onesided=None, nfft=None, fs=1, window=None, radialsum=False, radialavg=False):
    """periodogram(s; onesided, nfft, fs, window), src/periodograms.jl:393-417: the single-segment case; a matrix gives
    the two-dimensional / radial periodogram (:473-509)."""
    if not isinstance(s, DeviceArray):
        s = np.asarray(s)
    if s.ndim == 2:
        return _periodogram2(s, nfft, fs, radialsum, radialavg)
    if s.ndim != 1:
        raise ArgumentError("expected a vector or a matrix")
    cplx = s.dtype.kind == "c"
    onesided = (not cplx) if onesided is None else bool(onesided)
    if onesided and cplx:
        raise ArgumentError("cannot compute one-sided FFT of a complex signal")      # :396
    nfft = nextfastfft(s.size) if nfft is None else int(nfft)
    if nfft < s.size:
        raise DomainError("nfft must be >= n = length(s)")                           # :397
    if s.size == 0:
        raise ArgumentError("empty signal")
    win, norm2 = compute_window(window, s.size)
    if isinstance(s, DeviceArray):                                                  # device-resident signal, small result to the host
        if s.dtype != fftintype(s.dtype):
            raise ArgumentError("device signals must be Float32, Float64, ComplexF32 or ComplexF64")
        plan = _lib.SpecPlan(s.dtype, s.size, 0, nfft, onesided, win)
        dout = DeviceArray((plan.nout,), fftabs2type(s.dtype))
        plan.welch_dev(s.ptr, s.size, fs * norm2, dout.ptr, 0)
        out = dout.to_host()
        plan.close()
        return Periodogram(out, rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs))
    sig = _signal(s)
    plan = _lib.SpecPlan(sig.dtype, s.size, 0, nfft, onesided, win)
    out = np.empty(plan.nout, dtype=fftabs2type(sig.dtype))
    plan.welch(sig, fs * norm2, out)
    plan.close()
    return Periodogram(out, rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs))


def stft(s, n=None, noverlap=None, psdonly=False, onesided=None, nfft=None, fs=1, window=None):
    """stft(s, n, noverlap[, PSDOnly()]; onesided, nfft, fs, window), src/periodograms.jl:872-897.
    A 2-D `s` (len x nchan) is the batched extension: returns nout x k x nchan."""
    dev = isinstance(s, DeviceArray)
    if not dev:
        s = np.asarray(s)
    batched = s.ndim == 2
    if s.ndim not in (1, 2):
        raise ArgumentError("expected a vector (or a len x nchan matrix for the batched form)")
    length = s.shape[0]
    cplx = s.dtype.kind == "c"
    n = length >> 3 if n is None else int(n)
    noverlap = n >> 1 if noverlap is None else int(noverlap)
    onesided = (not cplx) if onesided is None else bool(onesided)
    nfft = nextfastfft(n) if nfft is None else int(nfft)
    if onesided and cplx:
        raise ArgumentError("cannot compute one-sided FFT of a complex signal")      # :876
    win, norm2 = compute_window(window, n)
    k = arraysplit_count(length, n, noverlap)
    if nfft < n:
        raise DomainError("nfft must be >= n")                                        # ArraySplit :45
    dt = fftintype(s.dtype)
    nout = nfft // 2 + 1 if onesided else nfft
    odt = fftabs2type(dt) if psdonly else fftouttype(dt)
    if dev:                                              # device pipeline form: the spectrogram matrix stays in HBM
        nchan = s.shape[1] if batched else 1
        dout = DeviceArray((nout, k, nchan) if batched else (nout, k), odt)
        if k > 0 and nchan > 0:
            plan = _lib.SpecPlan(dt, n, noverlap, nfft, onesided, win)
            plan.stft_dev(s.ptr, length, nchan, fs * norm2, psdonly, dout.ptr, 0)
            from .device import sync
            sync()
            plan.close()
        return dout
    sig = np.asfortranarray(s.reshape(length, -1), dtype=dt)
    nchan = sig.shape[1]
    out = np.zeros((nout, k, nchan), dtype=odt, order="F")
    if k > 0 and nchan > 0:
        plan = _lib.SpecPlan(dt, n, noverlap, nfft, onesided, win)
        plan.stft(sig, length, nchan, fs * norm2, psdonly, out)
        plan.close()
    return out if batched else out[:, :, 0]


def spectrogram(s, n=None, noverlap=None, onesided=None, nfft=None, fs=1, window=None):
    """spectrogram(s, n, noverlap; onesided, nfft, fs, window), src/periodograms.jl:828-837.
    A 2-D `s` (len x nchan) is the batched extension: power is nout x k x nchan."""
    if not isinstance(s, DeviceArray):
        s = np.asarray(s)
    length = s.shape[0]
    cplx = s.dtype.kind == "c"
    n = length >> 3 if n is None else int(n)
    noverlap = n >> 1 if noverlap is None else int(noverlap)
    onesided = (not cplx) if onesided is None else bool(onesided)
    nfft = nextfastfft(n) if nfft is None else int(nfft)
    out = stft(s, n, noverlap, psdonly=True, onesided=onesided, nfft=nfft, fs=fs, window=window)
    k = out.shape[1]
    t = (n / 2 + (n - noverlap) * np.arange(k, dtype=np.float64)) / fs               # :835
    return Spectrogram(out, rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs), t)
