"""Oracle: FIR application, fftfilt, polyphase resampling (reference src/Filters/filt.jl,
src/Filters/stream_filt.jl, src/Filters/design.jl). TEST INFRASTRUCTURE ONLY."""
import math
from fractions import Fraction

import numpy as np
import scipy.fft as sfft

from . import dspbase
from .dspbase import SMALL_FILT_CUTOFF, optimalfftfiltlength, promote
from .windows import kaiser


# --------------------------------------------------------------------------- fftfilt / filt(b, x)

def fftfilt(b, x, nfft=None, f64=False):
    """fftfilt / _fftfilt!, src/Filters/filt.jl:458-521: real overlap-save along dim 0, every column.
    Output has the shape of x.  Only Real b and Real x (:458-459)."""
    b = np.asarray(b)
    x = np.asarray(x)
    if np.iscomplexobj(b) or np.iscomplexobj(x):
        raise TypeError("fftfilt is defined for Real taps and Real signals only")
    W = promote(b.dtype, x.dtype)
    if not np.issubdtype(W, np.floating):
        W = np.dtype(np.float64)
    if f64:
        W = np.dtype(np.float64)
    nb = len(b)
    nx = x.shape[0]
    if nfft is None:
        nfft = optimalfftfiltlength(nb, x.size)  # note: length(x), :459
    x2 = x.reshape(nx, -1)
    out = np.zeros(x2.shape, dtype=W)
    L = min(nx, nfft - (nb - 1))                 # :490
    tmp1 = np.zeros(nfft, dtype=W)
    tmp1[:nb] = (b.astype(W) / W.type(nfft))     # :498 b ./ normfactor in W
    filterft = sfft.rfft(tmp1)
    for col in range(x2.shape[1]):
        off = 1
        while off <= nx:                          # :504 off = 1:L:nx
            npadbefore = max(0, nb - off)
            xstart = off - nb + npadbefore + 1
            n = min(nfft - npadbefore, nx - xstart + 1)
            tmp1 = np.zeros(nfft, dtype=W)
            tmp1[npadbefore: npadbefore + n] = x2[xstart - 1: xstart - 1 + n, col]
            tmp2 = sfft.rfft(tmp1) * filterft
            y = sfft.irfft(tmp2, n=nfft, norm="forward").astype(W)   # brfft (unnormalised)
            m = min(L, nx - off + 1)
            out[off - 1: off - 1 + m, col] = y[nb - 1: nb - 1 + m]
            off += L
    return out.reshape(x.shape)


def tdfilt(h, x, f64=False):
    """tdfilt, src/Filters/filt.jl:431-443 -> filt(h, one(H), x)."""
    return dspbase.filt(h, np.ones(1, dtype=np.asarray(h).dtype), x, f64=f64)


def filt(b, x, f64=False):
    """filt(b, x) algorithm chooser, src/Filters/filt.jl:525-555."""
    b = np.asarray(b)
    x = np.asarray(x)
    real = not (np.iscomplexobj(b) or np.iscomplexobj(x))
    if real and len(b) > SMALL_FILT_CUTOFF:       # :544-548
        nfft = optimalfftfiltlength(len(b), x.shape[0])
        return fftfilt(b, x, nfft, f64=f64)
    return tdfilt(b, x, f64=f64)                   # :549-555


# --------------------------------------------------------------------------- default resampling taps

def kaiserord(transitionwidth, attenuation=60):
    """src/Filters/design.jl:547-559."""
    n = math.ceil((attenuation - 7.95) / (math.pi * 2.285 * transitionwidth)) + 1
    if attenuation > 50:
        beta = 0.1102 * (attenuation - 8.7)
    elif attenuation >= 21:
        beta = 0.5842 * (attenuation - 21) ** 0.4 + 0.07886 * (attenuation - 21)
    else:
        beta = 0.0
    return n, beta / math.pi


def _sinc(t):
    return np.sinc(t)  # sin(pi t)/(pi t), same definition as Julia's sinc


def resample_filter(rate, rel_bw=1.0, attenuation=60):
    """resample_filter(rate::Union{Integer,Rational}), src/Filters/design.jl:694-720
    (+ lowpass FIRWindow prototype :598-602, scaling :642, :669-674)."""
    rate = Fraction(rate)
    nphi = rate.numerator
    dec = rate.denominator
    f_nyq = min(1 / nphi, 1 / dec)
    cutoff = f_nyq * rel_bw
    trans_width = cutoff * 0.2
    hlen, alpha = kaiserord(trans_width, attenuation)
    hlen = nphi * math.ceil(hlen / nphi)
    if hlen % 2 == 0:
        hlen += 1
    k = np.arange(1, hlen + 1, dtype=np.float64)
    w = cutoff  # normalize_freq(w, fs=2) = w
    coefs = w * _sinc(w * (k - (hlen + 1) / 2))
    coefs = coefs * kaiser(hlen, alpha)
    coefs = coefs * (1 / np.sum(coefs))
    return coefs * nphi


# --------------------------------------------------------------------------- polyphase machinery

def taps2pfb(h, nphi):
    """src/Filters/stream_filt.jl:294-307: tapsPerphi x Nphi, each column reversed (flipped up/down)."""
    h = np.asarray(h)
    hlen = len(h)
    tpp = math.ceil(hlen / nphi)
    hp = np.zeros(tpp * nphi, dtype=h.dtype)
    hp[:hlen] = h
    return hp.reshape(tpp, nphi)[::-1, :].copy()


def _round_half_even(v):
    return int(np.round(v))  # numpy rounds half to even, like Julia's round(Int, x)


def outputlength(inputlength, ratio, initial_phi):
    """src/Filters/stream_filt.jl:317-322."""
    ratio = Fraction(ratio)
    return math.ceil((inputlength * ratio.numerator - initial_phi + 1) / ratio.denominator)


def inputlength(outputlength_, ratio, initial_phi, round_up=False):
    """src/Filters/stream_filt.jl:358-364 (RoundDown default / RoundUp)."""
    ratio = Fraction(ratio)
    d = ratio.denominator if round_up else 1
    v = Fraction(outputlength_ * ratio.denominator + initial_phi - d, ratio.numerator)
    return math.ceil(v) if round_up else math.floor(v)


class FIRFilterState:
    """Literal stateful restatement of FIRFilter{FIRRational|FIRInterpolator|FIRDecimator|FIRStandard}
    (src/Filters/stream_filt.jl:137-178 ctor, :223-229 setphase!, :409-515, :522-560 filt! loops).
    Pure-Python loops: small cases only.  1-based indices are kept to mirror the reference."""

    def __init__(self, h, ratio=1):
        self.ratio = Fraction(ratio)
        self.h = np.asarray(h)
        I, D = self.ratio.numerator, self.ratio.denominator
        self.I, self.D = I, D
        self.hlen = len(self.h)
        if self.ratio == 1:
            self.kind = "standard"
            self.hrev = self.h[::-1].copy()
            self.history_len = self.hlen - 1
        elif D == 1:
            self.kind = "interp"
        elif I == 1:
            self.kind = "decim"
            self.hrev = self.h[::-1].copy()
            self.history_len = self.hlen - 1
        else:
            self.kind = "rational"
        if self.kind in ("interp", "rational"):
            self.pfb = taps2pfb(self.h, I)
            self.tpp = self.pfb.shape[0]
            self.history_len = self.tpp - 1
            self.phi_step = D % I
        self.phi_idx = 1
        self.input_deficit = 1
        self.history = None

    def timedelay(self):
        """:400-403."""
        if self.kind in ("interp", "rational"):
            return (self.hlen - 1) / (2 * self.I)
        return (self.hlen - 1) / 2

    def setphase(self, phi):
        """:216-229."""
        if self.kind in ("decim", "standard"):
            self.input_deficit += _round_half_even(phi)
        else:
            q, r = divmod(_round_half_even(phi * self.I), self.I)
            self.input_deficit += q
            self.phi_idx = r + 1

    def outputlength(self, inlen):
        """:324-338."""
        if self.kind == "standard":
            return inlen
        if self.kind == "interp":
            return outputlength(inlen - self.input_deficit + 1, self.I, self.phi_idx)
        if self.kind == "decim":
            return outputlength(inlen - self.input_deficit + 1, Fraction(1, self.D), 1)
        return outputlength(inlen - self.input_deficit + 1, self.ratio, self.phi_idx)

    def inputlength(self, outlen, round_up=False):
        """:366-383."""
        if self.kind == "standard":
            return outlen
        if self.kind == "interp":
            v = inputlength(outlen, self.I, self.phi_idx, round_up)
        elif self.kind == "decim":
            v = inputlength(outlen, Fraction(1, self.D), 1, round_up)
        else:
            v = inputlength(outlen, self.ratio, self.phi_idx, round_up)
        return v + self.input_deficit - 1

    def _dot(self, col, x, hist, last):
        """unsafe_dot(pfb, phi, [history,] x, last) -- src/util.jl:225-255 (1-based `last`)."""
        n = len(col)
        if last >= n:
            seg = x[last - n: last]
        else:
            seg = np.concatenate([hist[len(hist) - (n - last):], x[:last]])
        T = promote(col.dtype, x.dtype)
        return np.sum(col.astype(T) * seg.astype(T))

    def filt(self, x):
        x = np.asarray(x)
        T = promote(self.h.dtype, x.dtype)
        if self.history is None:
            self.history = np.zeros(self.history_len, dtype=x.dtype)
        hist = self.history
        xlen = len(x)
        out = []
        if self.kind == "standard":                                   # :409-428
            for i in range(1, xlen + 1):
                out.append(self._dot(self.hrev, x, hist, i))
        else:
            if xlen < self.input_deficit:                              # :484-488
                self.history = self._shiftin(hist, x)
                self.input_deficit -= xlen
                return np.zeros(0, dtype=T)
            idx = self.input_deficit
            while idx <= xlen:
                if self.kind == "decim":                               # :541-554
                    out.append(self._dot(self.hrev, x, hist, idx))
                    idx += self.D
                elif self.kind == "interp":                            # :448-461
                    out.append(self._dot(self.pfb[:, self.phi_idx - 1], x, hist, idx))
                    if self.phi_idx == self.I:
                        self.phi_idx, idx = 1, idx + 1
                    else:
                        self.phi_idx += 1
                else:                                                   # :496-509
                    out.append(self._dot(self.pfb[:, self.phi_idx - 1], x, hist, idx))
                    idx += (self.phi_idx + self.D - 1) // self.I
                    p = self.phi_idx + self.phi_step
                    self.phi_idx = p - self.I if p > self.I else p
            self.input_deficit = (idx - xlen) if self.kind != "interp" else 1
        self.history = self._shiftin(hist, x)
        return np.asarray(out, dtype=T)

    @staticmethod
    def _shiftin(a, b):
        """src/util.jl:299-314."""
        if len(a) == 0:
            return a
        return np.concatenate([a, b.astype(a.dtype)])[-len(a):]


def resample_literal(x, rate, h=None):
    """resample(x, rate[, h]) through the literal stateful loops (small inputs):
    src/Filters/stream_filt.jl:688-725."""
    x = np.asarray(x)
    rate = Fraction(rate)
    if h is None:
        h = resample_filter(rate)
    sf = FIRFilterState(h, rate)
    sf.setphase(sf.timedelay())                                          # undelay! :706-714
    outlen = math.ceil(len(x) * rate)                                    # :698
    xpad = np.zeros(sf.inputlength(outlen, round_up=True), dtype=x.dtype)  # :699
    xpad[:len(x)] = x[:len(xpad)] if len(xpad) < len(x) else x
    y = sf.filt(xpad)
    assert len(y) >= outlen, "Resample output shorter than expected."    # :722
    return y[:outlen]


def resample(x, rate, h=None, f64=False):
    """Vectorised closed form of resample for Integer / Rational rates,
    equivalent to `resample_literal` (checked in tests).  Output eltype promote_type(eltype(h), eltype(x))
    (src/Filters/stream_filt.jl:654).  Accumulation order matches unsafe_dot: taps in pfb-column order
    (oldest input sample first), plain multiply-add in the output dtype (or double when f64)."""
    x = np.asarray(x)
    rate = Fraction(rate)
    if h is None:
        h = resample_filter(rate)
    h = np.asarray(h)
    I, D = rate.numerator, rate.denominator
    T = promote(h.dtype, x.dtype)
    if f64:
        T = np.dtype(np.complex128 if np.issubdtype(T, np.complexfloating) else np.float64)
    hlen = len(h)
    tpp = math.ceil(hlen / I)
    hp = np.zeros(tpp * I, dtype=h.dtype)
    hp[:hlen] = h
    if rate == 1:
        n0, phi0 = 0, 0
        # FIRStandard: setphase! only throws samples away (round(tau)), :216-221 -- handled below
    tau = (hlen - 1) / (2 * I) if (I != 1) else (hlen - 1) / 2
    if I == 1:
        n0, phi0 = _round_half_even(tau), 0
    else:
        n0, phi0 = divmod(_round_half_even(tau * I), I)
    outlen = math.ceil(len(x) * rate)
    j = np.arange(outlen, dtype=np.int64)
    p = phi0 + j * D
    n = n0 + p // I
    phi = p % I
    xp = np.concatenate([np.zeros(tpp - 1, dtype=x.dtype), x, np.zeros(tpp + n0 + 2, dtype=x.dtype)]).astype(T)
    nmax = len(xp)
    acc = np.zeros(outlen, dtype=T)
    hpT = hp.astype(T)
    # pfb column order: row r (0-based, top = latest-index tap reversed) multiplies x[n - (tpp-1) + r]
    for r in range(tpp):
        t = tpp - 1 - r
        idx = n - t + (tpp - 1)
        idx = np.minimum(idx, nmax - 1)
        acc = (acc + hpT[phi + t * I] * xp[idx]).astype(T)
    return acc


# --------------------------------------------------------------------------- arbitrary-rate resampling

def resample_filter_arb(rate, nphi=32, rel_bw=1.0, attenuation=60):
    """resample_filter(rate::AbstractFloat, Nphi, rel_bw, attenuation), src/Filters/design.jl:683-686, 701-720."""
    f_nyq = 1.0 / nphi if rate >= 1.0 else rate / nphi
    cutoff = f_nyq * rel_bw
    hlen, alpha = kaiserord(cutoff * 0.2, attenuation)
    hlen = nphi * math.ceil(hlen / nphi)
    if hlen % 2 == 0:
        hlen += 1
    k = np.arange(1, hlen + 1, dtype=np.float64)
    coefs = cutoff * _sinc(cutoff * (k - (hlen + 1) / 2)) * kaiser(hlen, alpha)
    coefs = coefs * (1 / np.sum(coefs))
    return coefs * nphi


class FIRArbitraryState:
    """Literal restatement of FIRFilter{FIRArbitrary} (src/Filters/stream_filt.jl:92-134 ctor, :231-239 setphase!,
    :340-342 outputlength, :385-389 inputlength, :567-625 update! / filt!), including the serial Float64 phase accumulator.
    Pure-Python loop: small cases only."""

    def __init__(self, h, rate, nphi=32):
        self.h = np.asarray(h)
        self.rate = float(rate)
        self.nphi = int(nphi)
        dh = np.concatenate([np.diff(self.h), np.zeros(1, dtype=self.h.dtype)])
        self.pfb = taps2pfb(self.h, self.nphi)
        self.dpfb = taps2pfb(dh, self.nphi)
        self.tpp = self.pfb.shape[0]
        self.history_len = self.tpp - 1
        self.delta = self.nphi / self.rate
        self.hlen = len(self.h)
        self.acc, self.phi_idx, self.alpha, self.input_deficit, self.history = 0.0, 1, 0.0, 1, None

    def timedelay(self):
        return (self.hlen - 1) / (2 * self.nphi)

    def setphase(self, phi):
        frac, whole = math.modf(phi)
        self.input_deficit += _round_half_even(whole)
        self.acc = frac * self.nphi
        self.phi_idx = 1 + math.floor(self.acc)
        self.alpha = math.modf(self.acc)[0]

    def outputlength(self, inlen):
        return math.ceil((inlen - self.input_deficit + 1) * self.rate - self.acc / self.delta)

    def inputlength(self, outlen, round_up=False):
        d = 1 if round_up else 0
        return math.floor((outlen - d + self.acc / self.delta) / self.rate) + d + self.input_deficit - 1

    def _update(self):
        self.acc += self.delta
        x_adv = 0
        if self.acc >= self.nphi:
            q = math.floor(self.acc / self.nphi)            # divrem for positive operands
            self.acc = math.fmod(self.acc, self.nphi)
            x_adv = int(q)
        self.alpha, foffset = math.modf(self.acc)
        self.phi_idx = 1 + int(foffset)
        return x_adv

    def filt(self, x):
        x = np.asarray(x)
        T = promote(self.h.dtype, x.dtype)
        if self.history is None:
            self.history = np.zeros(self.history_len, dtype=x.dtype)
        hist, xlen, out = self.history, len(x), []
        if xlen < self.input_deficit:
            self.history = FIRFilterState._shiftin(hist, x)
            self.input_deficit -= xlen
            return np.zeros(0, dtype=T)
        idx = self.input_deficit
        Tw = np.dtype(np.complex128 if np.issubdtype(T, np.complexfloating) else np.float64)
        while idx <= xlen:
            lo = FIRFilterState._dot(self, self.pfb[:, self.phi_idx - 1], x, hist, idx)
            up = FIRFilterState._dot(self, self.dpfb[:, self.phi_idx - 1], x, hist, idx)
            out.append(np.asarray(Tw.type(up) * self.alpha + Tw.type(lo)).astype(T))      # muladd(yUpper, alpha::Float64, yLower)
            idx += self._update()
        self.input_deficit = idx - xlen
        self.history = FIRFilterState._shiftin(hist, x)
        return np.asarray(out, dtype=T)


def resample_arb_literal(x, rate, h=None, nphi=32):
    """resample(x, rate::AbstractFloat[, h, Nphi]), src/Filters/stream_filt.jl:692-725 through the literal loop."""
    x = np.asarray(x)
    if h is None:
        h = resample_filter_arb(rate, nphi)
    sf = FIRArbitraryState(h, rate, nphi)
    sf.setphase(sf.timedelay())
    outlen = math.ceil(len(x) * rate)
    xpad = np.zeros(sf.inputlength(outlen, round_up=True), dtype=x.dtype)
    m = min(len(x), len(xpad))
    xpad[:m] = x[:m]
    y = sf.filt(xpad)
    assert len(y) >= outlen, "Resample output shorter than expected."
    return y[:outlen]
