"""Oracle: periodogram / Welch / spectrogram / STFT (reference src/periodograms.jl). TEST INFRASTRUCTURE ONLY."""
import numpy as np
import scipy.fft as sfft

from .util import fftabs2type, fftfreq, fftintype, fftouttype, nextfastfft, rfftfreq


class DomainError(ValueError):
    pass


class DimensionMismatch(ValueError):
    pass


# --------------------------------------------------------------------------- segmenting / window

def arraysplit_count(length, n, noverlap):
    """src/periodograms.jl:49-50."""
    return (length - n) // (n - noverlap) + 1 if length >= n else 0


def arraysplit(s, n, noverlap, nfft=None, window=None, f64=False):
    """ArraySplit, src/periodograms.jl:32-73: returns the k x nfft matrix of (windowed, zero-padded)
    segments.  Buffer eltype fftintype(eltype(s)) (:55); window*sample formed in the promoted type
    and rounded to the buffer eltype (:66)."""
    s = np.asarray(s)
    nfft = n if nfft is None else nfft
    if not (0 <= noverlap < n):
        raise DomainError("noverlap must be between zero and n")   # :44
    if nfft < n:
        raise DomainError("nfft must be >= n")                      # :45
    S = fftintype(s.dtype)
    if f64:
        S = np.dtype(np.complex128 if np.issubdtype(S, np.complexfloating) else np.float64)
    k = arraysplit_count(len(s), n, noverlap)
    hop = n - noverlap
    out = np.zeros((k, nfft), dtype=S)
    if k == 0:
        return out
    idx = (np.arange(k) * hop)[:, None] + np.arange(n)[None, :]
    seg = s[idx]
    if window is None:
        out[:, :n] = seg.astype(S)
    else:
        w = np.asarray(window)
        P = np.result_type(s.dtype, w.dtype)
        out[:, :n] = (seg.astype(P) * w.astype(P)[None, :]).astype(S)
    return out


def compute_window(window, n):
    """src/periodograms.jl:248-257."""
    if window is None:
        return None, n
    if callable(window):
        win = np.asarray(window(n), dtype=np.float64)
        return win, float(np.sum(np.abs(win) ** 2))
    win = np.asarray(window)
    if len(win) != n:
        raise DimensionMismatch("length of window must match input")
    return win, float(np.sum(np.abs(win) ** 2))


# --------------------------------------------------------------------------- power scaling

def fft2pow_terms(s_fft, nfft, r, onesided, T):
    """Per-segment PSD terms |X|^2 * m of fft2pow!, src/periodograms.jl:142-172, for a (k x n) batch of
    spectra -> (k x nout) in eltype T.  m1 = 1/r, m2 = 2/r are computed in Float64 and converted to T
    (:143,146); abs2 is evaluated in T."""
    T = np.dtype(T)
    m1 = T.type(1 / r)
    m2 = T.type(2 / r)
    p = (s_fft.real.astype(T) ** 2 + s_fft.imag.astype(T) ** 2).astype(T)
    n = s_fft.shape[1]
    if onesided:
        m = np.full(n, m2, dtype=T)
        m[0] = m1
        m[n - 1] = m1 if nfft % 2 == 0 else m2
        return p * m[None, :]
    if n == nfft:
        return p * m1
    terms = np.zeros((p.shape[0], nfft), dtype=T)       # real FFT -> two-sided, :157-169
    terms[:, :n] = p * m1
    i = np.arange(2, n)                                  # 1-based i = 2:n-1
    terms[:, nfft - i + 1] = p[:, i - 1] * m1
    if nfft % 2 == 1:
        terms[:, n] = p[:, n - 1] * m1
    return terms


def fft2pow_acc(out, s_fft, nfft, r, onesided, sequential=False):
    """Accumulate fft2pow! terms of a batch into `out`.  `sequential=True` adds segment by segment with
    one rounding per term in eltype(out) (the reference's order, muladd ~ fma); otherwise numpy sums
    the batch in eltype(out) (pairwise), which is at least as accurate."""
    T = out.dtype
    terms = fft2pow_terms(s_fft, nfft, r, onesided, T)
    if sequential:
        W = np.float64 if T == np.float32 else np.longdouble
        for row in terms:
            out[:] = (out.astype(W) + row.astype(W)).astype(T)
    else:
        out += terms.sum(axis=0, dtype=T)
    return out


def fft2oneortwosided(s_fft, nfft, onesided):
    """fft2oneortwosided!, src/periodograms.jl:234-244 on a (k x n) batch -> (k x nout)."""
    n = s_fft.shape[1]
    if onesided or n == nfft:
        return s_fft.copy()
    out = np.zeros((s_fft.shape[0], nfft), dtype=s_fft.dtype)
    out[:, :n] = s_fft
    i = np.arange(2, n - (1 if nfft % 2 == 0 else 0) + 1)   # 1-based 2 : n - iseven(nfft)
    out[:, nfft - i + 1] = np.conj(s_fft[:, i - 1])
    return out


def _forward(segs):
    """forward_plan, src/periodograms.jl:511-514: rfft for real buffers, fft for complex."""
    if np.iscomplexobj(segs):
        return sfft.fft(segs, axis=1)
    return sfft.rfft(segs, axis=1)


# --------------------------------------------------------------------------- public entry points

def periodogram(s, onesided=None, nfft=None, fs=1.0, window=None, f64=False):
    """periodogram (1-D), src/periodograms.jl:393-417.  Returns (power, freq)."""
    s = np.asarray(s)
    cplx = np.iscomplexobj(s)
    onesided = (not cplx) if onesided is None else onesided
    if onesided and cplx:
        raise ValueError("cannot compute one-sided FFT of a complex signal")   # ArgumentError :396
    nfft = nextfastfft(len(s)) if nfft is None else nfft
    if nfft < len(s):
        raise DomainError("nfft must be >= n = length(s)")                      # :397
    win, norm2 = compute_window(window, len(s))
    segs = arraysplit(s, len(s), 0, nfft, win, f64=f64) if len(s) > 0 else np.zeros((1, nfft))
    X = _forward(segs)
    T = np.dtype(np.float64) if f64 else fftabs2type(s.dtype)
    out = np.zeros(nfft // 2 + 1 if onesided else nfft, dtype=T)
    fft2pow_acc(out, X, nfft, fs * norm2, onesided)
    return out, (rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs))


def welch_pgram(s, n=None, noverlap=None, onesided=None, nfft=None, fs=1.0, window=None,
                f64=False, sequential=False):
    """welch_pgram, src/periodograms.jl:647-649, 560-576, 746-759.  Returns (power, freq).
    (`window=None` is the explicit `window=nothing`; the deprecated default is also `nothing`, :582-587.)"""
    s = np.asarray(s)
    cplx = np.iscomplexobj(s)
    n = len(s) >> 3 if n is None else n
    noverlap = n >> 1 if noverlap is None else noverlap
    onesided = (not cplx) if onesided is None else onesided
    nfft = nextfastfft(n) if nfft is None else nfft
    if onesided and cplx:
        raise ValueError("cannot compute one-sided FFT of a complex signal")   # :564
    if nfft < n:
        raise DomainError("nfft must be >= n")                                  # :565
    win, norm2 = compute_window(window, n)
    r0 = fs * norm2                                                             # :568
    segs = arraysplit(s, n, noverlap, nfft, win, f64=f64)
    k = segs.shape[0]
    T = np.dtype(np.float64) if f64 else fftabs2type(s.dtype)
    out = np.zeros(nfft // 2 + 1 if onesided else nfft, dtype=T)                # fill!(out, 0) :747
    r = k * r0                                                                  # :751
    if k > 0:
        # process in slabs to bound memory
        step = max(1, (1 << 24) // max(nfft, 1))
        for i in range(0, k, step):
            fft2pow_acc(out, _forward(segs[i:i + step]), nfft, r, onesided, sequential=sequential)
    return out, (rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs))


def stft(s, n=None, noverlap=None, psdonly=False, onesided=None, nfft=None, fs=1.0, window=None, f64=False):
    """stft, src/periodograms.jl:872-897.  Returns the (nout x k) matrix (column = segment):
    raw spectra (fftouttype) or, with psdonly, PSD columns (fftabs2type) scaled by r = fs*norm2."""
    s = np.asarray(s)
    cplx = np.iscomplexobj(s)
    n = len(s) >> 3 if n is None else n
    noverlap = n >> 1 if noverlap is None else noverlap
    onesided = (not cplx) if onesided is None else onesided
    nfft = nextfastfft(n) if nfft is None else nfft
    if onesided and cplx:
        raise ValueError("cannot compute one-sided FFT of a complex signal")   # :876
    win, norm2 = compute_window(window, n)
    segs = arraysplit(s, n, noverlap, nfft, win, f64=f64)
    k = segs.shape[0]
    nout = nfft // 2 + 1 if onesided else nfft
    r = fs * norm2
    if psdonly:
        T = np.dtype(np.float64) if f64 else fftabs2type(s.dtype)
        out = np.zeros((nout, k), dtype=T)
    else:
        T = np.dtype(np.complex128) if f64 else fftouttype(s.dtype)
        out = np.zeros((nout, k), dtype=T)
    if k == 0:
        return out
    X = _forward(segs)
    if psdonly:
        out[:, :] = fft2pow_terms(X, nfft, r, onesided, T).T
    else:
        out[:, :] = fft2oneortwosided(X, nfft, onesided).T.astype(T)
    return out


def spectrogram(s, n=None, noverlap=None, onesided=None, nfft=None, fs=1.0, window=None, f64=False):
    """spectrogram, src/periodograms.jl:828-837.  Returns (power nout x k, freq, time)."""
    s = np.asarray(s)
    cplx = np.iscomplexobj(s)
    n = len(s) >> 3 if n is None else n
    noverlap = n >> 1 if noverlap is None else noverlap
    onesided = (not cplx) if onesided is None else onesided
    nfft = nextfastfft(n) if nfft is None else nfft
    out = stft(s, n, noverlap, psdonly=True, onesided=onesided, nfft=nfft, fs=fs, window=window, f64=f64)
    k = out.shape[1]
    time = (n / 2 + (n - noverlap) * np.arange(k)) / fs                          # :835
    return out, (rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs)), time


# --------------------------------------------------------------------------- multitaper

def mt_pgram(s, onesided=None, nfft=None, fs=1.0, nw=4, ntapers=None, window=None, taper_weights=None, f64=False):
    """mt_pgram, src/multitaper.jl:178-242 (MTConfig :117-141, mt_fft_tapered! :149-159): sum over tapers of
    fft2pow!(FFT(window[:, t] .* s), r[t]) with r = fs ./ weights (dpss default, unit-norm tapers) or
    fs .* sum(abs2, window; dims=1) ./ weights (explicit window).  Returns (power, freq)."""
    import math
    from .windows import dpss
    s = np.asarray(s)
    cplx = np.iscomplexobj(s)
    onesided = (not cplx) if onesided is None else onesided
    n = len(s)
    nfft = nextfastfft(n) if nfft is None else nfft
    if onesided and cplx:
        raise ValueError("cannot compute one-sided FFT of a complex signal")
    if nfft < n:
        raise ValueError("Must have `nfft >= n_samples`")
    ntapers = (math.ceil(2 * nw) - 1) if ntapers is None else ntapers
    if window is None:
        win = dpss(n, nw, ntapers)
        norm2 = np.ones(ntapers)
    else:
        win = np.asarray(window, dtype=np.float64)
        ntapers = win.shape[1]
        norm2 = np.sum(np.abs(win) ** 2, axis=0)
    w = np.full(ntapers, 1.0 / ntapers) if taper_weights is None else np.asarray(taper_weights, dtype=np.float64)
    r = fs * norm2 / w
    S = fftintype(s.dtype)
    if f64:
        S = np.dtype(np.complex128 if cplx else np.float64)
    T = np.dtype(np.float64) if f64 else fftabs2type(s.dtype)
    out = np.zeros(nfft // 2 + 1 if onesided else nfft, dtype=T)
    for t in range(ntapers):
        buf = np.zeros((1, nfft), dtype=S)
        buf[0, :n] = (win[:, t] * s).astype(S)
        X = sfft.fft(buf, axis=1) if (cplx or not onesided) else sfft.rfft(buf, axis=1)
        fft2pow_acc(out, X, nfft, r[t], onesided)
    return out, (rfftfreq(nfft, fs) if onesided else fftfreq(nfft, fs))


def mt_spectrogram(s, n=None, n_overlap=None, fs=1.0, onesided=None, nfft=None, nw=4, ntapers=None, window=None, f64=False):
    """mt_spectrogram, src/multitaper.jl:262-404: one mt_pgram per segment (segments from arraysplit);
    default nfft = nextpow(2, n) (MTConfig default, :117).  Returns (power nout x k, freq, time)."""
    s = np.asarray(s)
    n = len(s) >> 3 if n is None else n
    n_overlap = n >> 1 if n_overlap is None else n_overlap
    if n <= n_overlap:
        raise ValueError("Need `samples_per_window > n_overlap_samples`")
    nfft = (1 << (n - 1).bit_length()) if nfft is None else nfft
    hop = n - n_overlap
    k = 0 if len(s) < n else (len(s) - n) // hop + 1
    cols = []
    f = None
    for i in range(k):
        p, f = mt_pgram(s[i * hop: i * hop + n], onesided=onesided, nfft=nfft, fs=fs, nw=nw, ntapers=ntapers, window=window, f64=f64)
        cols.append(p)
    if f is None:
        cplx = np.iscomplexobj(s)
        os_ = (not cplx) if onesided is None else onesided
        f = rfftfreq(nfft, fs) if os_ else fftfreq(nfft, fs)
    power = np.stack(cols, axis=1) if cols else np.zeros((len(f), 0))
    time = (n / 2 + hop * np.arange(k)) / fs
    return power, f, time


# ------------------------------------------------- multitaper cross spectra / coherence

def dpss_config(n_samples, nw=4, ntapers=None, keep_only_large_evals=False, weight_by_evals=False):
    """dpss_config, src/multitaper.jl:52-77 -> (window n x ntapers, taper_weights)."""
    from .windows import dpss, dpsseig
    ntapers = int(2 * nw - 1) if ntapers is None else ntapers
    window = dpss(n_samples, nw, ntapers)
    evals = None
    if keep_only_large_evals:
        evals = dpsseig(window, nw)
        keep = evals > 0.9
        window, evals = window[:, keep], evals[keep]
    if weight_by_evals:
        if evals is None:
            evals = dpsseig(window, nw)
        weights = evals / evals.sum()
    else:
        weights = np.full(window.shape[1], 1.0 / window.shape[1])
    return window, weights


def mt_cross_power_spectra(signal, fs=1.0, nfft=None, window=None, taper_weights=None, nw=4, ntapers=None, demean=False,
                           freq_range=None, f64=True):
    """mt_cross_power_spectra, src/multitaper.jl:470-603: signal is n_channels x n_samples (real, one-sided only, :411-416);
    x_mt[f, t, c] = rfft(window[:, t] .* signal[c, :], nfft) (:149-159, 589-593), DC (and Nyquist for even nfft) rows
    divided by sqrt(2) (:576-579), output[l, m, f] = sum_t (2 / r_t) x_mt[f, t, l] conj(x_mt[f, t, m]) (:595-616) over the
    frequencies strictly inside freq_range (:497-503).  Returns (power n_chan x n_chan x nf complex, freq)."""
    from .windows import dpss
    signal = np.asarray(signal)
    if np.iscomplexobj(signal):
        raise ValueError("Only real data is supported (with the default choice of `onesided=true`) for this operation.")
    n_chan, n = signal.shape
    nfft = (1 << (n - 1).bit_length()) if nfft is None else nfft             # nextpow(2, n_samples), :117
    T = np.dtype(np.float64) if f64 else fftintype(signal.dtype)
    if window is None:
        ntapers = int(2 * nw - 1) if ntapers is None else ntapers
        win = dpss(n, nw, ntapers)
        norm2 = np.ones(ntapers)
    else:
        win = np.asarray(window, dtype=np.float64)
        norm2 = np.sum(win * win, axis=0)
    K = win.shape[1]
    w = np.full(K, 1.0 / K) if taper_weights is None else np.asarray(taper_weights, dtype=np.float64)
    r = fs * norm2 / w
    sig = signal.astype(T)
    if demean:
        sig = sig - sig.mean(axis=1, keepdims=True).astype(T)
    freq = rfftfreq(nfft, fs)
    idx = np.arange(freq.size) if freq_range is None else np.flatnonzero((freq_range[0] < freq) & (freq < freq_range[-1]))
    x_mt = np.empty((nfft // 2 + 1, K, n_chan), dtype=np.complex128)
    for c in range(n_chan):
        for t in range(K):
            buf = np.zeros(nfft, dtype=T)
            buf[:n] = (win[:, t] * sig[c]).astype(T)
            x_mt[:, t, c] = np.fft.rfft(buf.astype(np.float64))
    x_mt[0] /= np.sqrt(2)
    if nfft % 2 == 0:
        x_mt[-1] /= np.sqrt(2)
    xs = x_mt[idx]                                                           # nf x K x C
    out = np.einsum("t,ftl,ftm->lmf", 2.0 / r, xs, np.conj(xs))
    return out.astype(np.complex128 if T == np.float64 else np.complex64), freq[idx]


def coherence_from_cs(cs):
    """coherence_from_cs!, src/multitaper.jl:672-693: |S_lm| / sqrt(real(S_ll S_mm)), unit diagonal."""
    n_chan = cs.shape[0]
    d = np.real(np.einsum("iif->if", cs))
    out = np.abs(cs) / np.sqrt(d[:, None, :] * d[None, :, :])
    low = np.tril(np.ones((n_chan, n_chan), dtype=bool), -1)[:, :, None]
    out = np.where(low, out, 0.0)
    out = out + np.transpose(out, (1, 0, 2))
    for i in range(n_chan):
        out[i, i, :] = 1.0
    return out.astype(np.float32 if cs.dtype == np.complex64 else np.float64)


def mt_coherence(signal, **kw):
    """mt_coherence, src/multitaper.jl:722-790 -> (coherence n_chan x n_chan x nf, freq)."""
    cs, freq = mt_cross_power_spectra(signal, **kw)
    return coherence_from_cs(cs), freq


# ------------------------------------------------- 2-D periodogram

def periodogram2(s, nfft=None, fs=1.0, radialsum=False, radialavg=False):
    """periodogram(s::AbstractMatrix; nfft, fs, radialsum, radialavg), src/periodograms.jl:473-509 with fft2pow2! (:175-181)
    and the literal fft2pow2radial! loop (:183-232).  Returns (power, freq1, freq2) or, for the radial forms, (power, freq)."""
    s = np.asarray(s)
    if nfft is None:
        nfft = tuple(nextfastfft(n) for n in s.shape)
    if not (s.shape[0] <= nfft[0] and s.shape[1] <= nfft[1]):
        raise ValueError("nfft must be >= size(s)")
    if not (s.shape[0] > 1 and s.shape[1] > 1):
        raise ValueError("dimensions of s must be > 1")
    if radialsum and radialavg:
        raise ValueError("radialsum and radialavg are mutually exclusive")
    T = fftabs2type(s.dtype)
    S = fftintype(s.dtype)
    norm2 = s.size
    r = fs * norm2
    inp = np.zeros(nfft, dtype=S)
    inp[:s.shape[0], :s.shape[1]] = s
    if not (radialsum or radialavg):
        X = np.fft.fft2(inp.astype(np.float64))
        return (np.abs(X) ** 2 * (1 / r)).astype(T), fftfreq(nfft[0], fs), fftfreq(nfft[1], fs)
    n1, n2 = nfft
    X = np.fft.fft2(inp.astype(np.float64))[: n1 // 2 + 1, :]            # rfft halves the first dimension
    nmin = min(n1, n2)
    n1max = (n1 >> 1) + 1
    kmax = (nmin >> 1) + 1
    out = np.zeros(kmax, dtype=T)
    wc = np.zeros(kmax, dtype=np.int64)
    m1, m2 = T.type(1 / r), T.type(2 / r)
    if n1 == nmin:
        c2, c1 = n1 / n2, 1.0
    else:
        c1, c2 = n2 / n1, 1.0
    P = (np.abs(X) ** 2).astype(T)

    def rnd(v):
        return int(np.rint(v))                                            # round(Int, x): ties to even

    for j in range(1, n2 + 1):
        kj1 = j - 1 if j <= (n2 >> 1) + 1 else -n2 + j - 1
        kj2 = (kj1 * c2) ** 2
        for i in range(1, n1max + 1):
            a = c1 * (i - 1)
            wavenum = rnd(np.sqrt(a * a + kj2)) + 1
            if wavenum <= kmax:
                if i == 1:
                    m, cnt = m1, 1
                elif i == n1max:
                    m, cnt = (m1, 1) if n1 % 2 == 0 else (m2, 2)
                else:
                    m, cnt = m2, 2
                out[wavenum - 1] = T.type(P[i - 1, j - 1] * m + out[wavenum - 1])
                wc[wavenum - 1] += cnt
    if radialavg:
        out = (out / wc).astype(T)
    return out, np.arange(kmax) * (fs / nmin)
