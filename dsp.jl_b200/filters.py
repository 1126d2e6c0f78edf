"""FIR application and polyphase resampling front ends (reference src/Filters/filt.jl:431-555,
src/Filters/stream_filt.jl, src/Filters/design.jl:547-559, 694-720), backed by libdspb200."""
import math
from fractions import Fraction

import numpy as np

from . import _lib
from .device import DeviceArray
from .dspbase import SMALL_FILT_CUTOFF, _cols, _gpu_dtype, _os_plan, _promote, filt_ as _filt_ba, optimalfftfiltlength
from .errors import ArgumentError, DomainError
from .windows import kaiser


# --------------------------------------------------------------------------------------------- tdfilt / fftfilt / filt(h, x)

def tdfilt(h, x):
    """tdfilt(h, x), src/Filters/filt.jl:431-433 -> filt(h, one(H), x)."""
    h = np.asarray(h)
    x = np.asarray(x)
    out = np.empty(x.shape, dtype=_gpu_dtype(_promote(h, x)), order="F")
    return tdfilt_(out, h, x)


def tdfilt_(out, h, x):
    """tdfilt!(out, h, x), src/Filters/filt.jl:441-443."""
    h = np.asarray(h)
    return _filt_ba(out, h, np.ones(1, dtype=h.dtype), x)


def _require_real(b, x):
    if np.iscomplexobj(b) or np.iscomplexobj(x):
        raise TypeError("fftfilt is defined for Real taps and Real signals only (src/Filters/filt.jl:458-459)")


def fftfilt(b, x, nfft=None):
    """fftfilt(b, x[, nfft]), src/Filters/filt.jl:458-461: real overlap-save along axis 0 of every column.
    nfft=None lets the library choose the block transform (the reference default is the CPU cost model
    optimalfftfiltlength); an explicit nfft is honoured."""
    b = np.asarray(b)
    if isinstance(x, DeviceArray):                       # device pipeline form: same-length overlap-save, stays in HBM
        if np.iscomplexobj(b) or x.dtype.kind == "c":
            raise TypeError("fftfilt is defined for Real taps and Real signals only (src/Filters/filt.jl:458-459)")
        nx = x.shape[0]
        out = DeviceArray(x.shape, x.dtype)
        if x.size:
            _os_plan(np.ascontiguousarray(b, dtype=x.dtype), nfft).exec_dev(x.ptr, nx, x.size // nx, out.ptr, nx, 0)
        return out
    x = np.asarray(x)
    _require_real(b, x)
    out = np.empty(x.shape, dtype=_gpu_dtype(_promote(b, x)), order="F")
    return _fftfilt(out, b, x, nfft)


def fftfilt_(out, b, x, nfft=None):
    """fftfilt!(out, b, x[, nfft]), src/Filters/filt.jl:468-476."""
    b = np.asarray(b)
    x = np.asarray(x)
    _require_real(b, x)
    if out.shape != x.shape:
        raise ArgumentError("out and x must be the same size")
    return _fftfilt(out, b, x, nfft)


def _fftfilt(out, b, x, nfft):
    """_fftfilt!, src/Filters/filt.jl:479-521 (the block loop runs on the GPU)."""
    if b.size == 0:
        raise ArgumentError("filter vector b must be non-empty")
    W = _gpu_dtype(_promote(b, x))
    if x.size == 0:
        return out
    xW, nx, ncols = _cols(x, W)
    bW = np.ascontiguousarray(b, dtype=W)
    if nfft is not None and nfft < b.size:
        raise ArgumentError("nfft must be >= length(b)")     # the reference leaves this unchecked (garbage result)
    res = np.empty((nx, ncols), dtype=W, order="F")
    _os_plan(bW, nfft).exec(xW, res, nx, ncols, nx)
    out[...] = res.reshape(x.shape)
    return out


def filt(b, x):
    """filt(b, x), src/Filters/filt.jl:445-446, 525-527."""
    b = np.asarray(b)
    x = np.asarray(x)
    out = np.empty(x.shape, dtype=_gpu_dtype(_promote(b, x)), order="F")
    return _filt_choose_alg(out, b, x)


def filt_(out, b, x):
    """filt!(out, b, x), src/Filters/filt.jl:530-533."""
    b = np.asarray(b)
    x = np.asarray(x)
    if out.shape != x.shape:
        raise ArgumentError("out must be the same size as x")
    return _filt_choose_alg(out, b, x)


def _filt_choose_alg(out, b, x):
    """filt_choose_alg!, src/Filters/filt.jl:537-555: Real x Real with nb > 66 -> overlap-save, else time domain."""
    real = not (np.iscomplexobj(b) or np.iscomplexobj(x))
    if real and b.size > SMALL_FILT_CUTOFF:
        return _fftfilt(out, b, x, None)
    return tdfilt_(out, b, x)


# --------------------------------------------------------------------------------------------- default taps

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


def _resample_filter(f_nyq, nphi, rel_bw, attenuation):
    """_resample_filter, src/Filters/design.jl:701-720: Kaiser-windowed sinc lowpass, length rounded up to an odd multiple
    of Nphi, DC gain Nphi."""
    cutoff = f_nyq * rel_bw
    hlen, alpha = kaiserord(cutoff * 0.2, attenuation)
    hlen = nphi * math.ceil(hlen / nphi)
    hlen += (hlen % 2 == 0)
    k = np.arange(1, hlen + 1, dtype=np.float64)
    h = cutoff * np.sinc(cutoff * (k - (hlen + 1) / 2)) * kaiser(hlen, alpha)   # :598-602, FIRWindow :669-674
    h *= 1 / h.sum()                                                              # scalefactor(Lowpass) :642
    return h * nphi                                                               # rmul!(h, Nphi) :719


def resample_filter(rate, *args):
    """resample_filter(rate::Union{Integer,Rational}, rel_bw=1.0, attenuation=60) (src/Filters/design.jl:694-699) and
    resample_filter(rate::AbstractFloat, Nphi=32, rel_bw=1.0, attenuation=60) (:683-686)."""
    if isinstance(rate, (float, np.floating)):
        nphi, rel_bw, attenuation = (list(args) + [32, 1.0, 60][len(args):])[:3]
        nphi = int(nphi)
        return _resample_filter(1.0 / nphi if rate >= 1.0 else rate / nphi, nphi, rel_bw, attenuation)
    rel_bw, attenuation = (list(args) + [1.0, 60][len(args):])[:2]
    rate = _as_ratio(rate)
    nphi, dec = rate.numerator, rate.denominator
    return _resample_filter(min(1 / nphi, 1 / dec), nphi, rel_bw, attenuation)


def _as_ratio(rate):
    if isinstance(rate, (float, np.floating)):
        raise ArgumentError("a floating-point rate selects the arbitrary-rate (FIRArbitrary) path; pass an int or "
                            "fractions.Fraction here")
    if isinstance(rate, str):
        rate = rate.replace("//", "/")
    return Fraction(rate)


def _round_half_even(v):
    return int(np.round(v))     # Julia's round(Int, x) is ties-to-even


def resample_phase(hlen, rate):
    """undelay! -> setphase!(timedelay): src/Filters/stream_filt.jl:216-229, 400-403, 706-714.
    Returns (n0, phi0): input samples skipped and 0-based start phase."""
    I, D = rate.numerator, rate.denominator
    if I == 1:                                    # FIRStandard / FIRDecimator: tau = (hLen-1)/2, whole samples only
        return _round_half_even((hlen - 1) / 2), 0
    tau = (hlen - 1) / (2 * I)
    q, r = divmod(_round_half_even(tau * I), I)
    return q, r


# --------------------------------------------------------------------------------------------- FIRFilter (stateful)

def outputlength(inputlength_, ratio, initial_phi):
    """outputlength(inputlength, ratio, initialphi), src/Filters/stream_filt.jl:317-322 (initialphi is 1-based)."""
    ratio = _as_ratio(ratio)
    return -((-(inputlength_ * ratio.numerator - initial_phi + 1)) // ratio.denominator)


def inputlength(outputlength_, ratio, initial_phi, round_up=False):
    """inputlength(outputlength, ratio, initialphi, RoundDown | RoundUp), src/Filters/stream_filt.jl:358-364."""
    ratio = _as_ratio(ratio)
    d = ratio.denominator if round_up else 1
    num = outputlength_ * ratio.denominator + initial_phi - d
    return -((-num) // ratio.numerator) if round_up else num // ratio.numerator


class FIRFilter:
    """FIRFilter(h, ratio=1): stateful single-rate / interpolating / decimating / rational polyphase FIR filter,
    src/Filters/stream_filt.jl:137-178 (kernels :8-78).  State carried across `filt` calls exactly as the
    reference: `history` (last historyLen input samples), `phi_idx` (1-based phase) and `input_deficit`
    (:476-515).  Each call runs the polyphase kernel on [history; x] through the closed form of the phase
    recurrence.

    FIRFilter(h, rate::float, nphases=32) is the arbitrary-rate filter (FIRArbitrary, :92-134, 193-205): polyphase bank
    plus derivative bank with linear interpolation between phases; state `phi_accumulator` / `input_deficit`.  The
    per-call phase sequence acc0 + j*delta is evaluated exactly (rational arithmetic on the host for the counts,
    double-double on the device), see resample.cu.  `h=None` designs the taps with resample_filter(rate, nphases)."""

    def __init__(self, h, ratio=1, nphases=32):
        if isinstance(ratio, (float, np.floating)):
            self._init_arbitrary(h, float(ratio), int(nphases))
            return
        self.ratio = _as_ratio(ratio)
        if self.ratio <= 0:
            raise ArgumentError("ratio must be positive")
        h = np.asarray(h)
        if h.ndim != 1 or h.size == 0:
            raise ArgumentError("h must be a non-empty vector")
        if np.iscomplexobj(h):
            raise NotImplementedError("complex taps are outside the B200 hot-path scope")
        self.h = np.ascontiguousarray(h, dtype=np.float32 if h.dtype == np.float32 else np.float64)
        self.interpolation, self.decimation = self.ratio.numerator, self.ratio.denominator
        self.hlen = self.h.size
        I, D = self.interpolation, self.decimation
        if self.ratio == 1:
            self.kind, self.history_len = "standard", self.hlen - 1
        elif D == 1:
            self.kind, self.history_len = "interpolator", -(-self.hlen // I) - 1
        elif I == 1:
            self.kind, self.history_len = "decimator", self.hlen - 1
        else:
            self.kind, self.history_len = "rational", -(-self.hlen // I) - 1
        self.taps_per_phase = -(-self.hlen // I)
        self._plans = {}
        self.reset()

    def _init_arbitrary(self, h, rate, nphases):
        if not rate > 0.0:
            raise DomainError("rate must be greater than 0")                                  # :194
        if h is None:
            h = resample_filter(rate, nphases)
        h = np.asarray(h)
        if h.ndim != 1 or h.size == 0:
            raise ArgumentError("h must be a non-empty vector")
        if np.iscomplexobj(h):
            raise NotImplementedError("complex taps are outside the B200 hot-path scope")
        self.h = np.ascontiguousarray(h, dtype=np.float32 if h.dtype == np.float32 else np.float64)
        self.kind, self.rate, self.nphases = "arbitrary", rate, nphases
        self.ratio = rate
        self.hlen = self.h.size
        self.taps_per_phase = -(-self.hlen // nphases)
        self.history_len = self.taps_per_phase - 1
        self.delta = nphases / rate                                                           # :115
        self._plans = {}
        self.reset()

    def reset(self):
        """reset!, src/Filters/stream_filt.jl:247-276."""
        self.phi_idx = 1
        self.input_deficit = 1
        self.history = None
        self.phi_accumulator = 0.0
        return self

    @property
    def alpha(self):
        return math.modf(self.phi_accumulator)[0]

    def timedelay(self):
        """timedelay, src/Filters/stream_filt.jl:400-403."""
        if self.kind == "arbitrary":
            return (self.hlen - 1) / (2 * self.nphases)
        if self.kind in ("rational", "interpolator"):
            return (self.hlen - 1) / (2 * self.interpolation)
        return (self.hlen - 1) / 2

    def setphase(self, phi):
        """setphase!, src/Filters/stream_filt.jl:216-229."""
        if phi < 0:
            raise DomainError("phi must be >= 0")
        if self.kind == "arbitrary":                                                          # :231-239
            frac, whole = math.modf(phi)
            self.input_deficit += _round_half_even(whole)
            self.phi_accumulator = frac * self.nphases
            self.phi_idx = 1 + math.floor(self.phi_accumulator)
        elif self.kind in ("rational", "interpolator"):
            q, r = divmod(_round_half_even(phi * self.interpolation), self.interpolation)
            self.input_deficit += q
            self.phi_idx = r + 1
        else:
            self.input_deficit += _round_half_even(phi)

    def outputlength(self, inlen):
        """outputlength(::FIRFilter, inputlength), src/Filters/stream_filt.jl:324-342."""
        if self.kind == "standard":
            return inlen
        if self.kind == "arbitrary":                                                          # :340-342
            return math.ceil((inlen - self.input_deficit + 1) * self.rate - self.phi_accumulator / self.delta)
        return outputlength(inlen - self.input_deficit + 1, self.ratio, self.phi_idx if self.kind != "decimator" else 1)

    def inputlength(self, outlen, round_up=False):
        """inputlength(::FIRFilter, outputlength, r), src/Filters/stream_filt.jl:366-398."""
        if self.kind == "standard":
            return outlen
        if self.kind == "arbitrary":                                                          # :385-389
            d = 1 if round_up else 0
            return math.floor((outlen - d + self.phi_accumulator / self.delta) / self.rate) + d + self.input_deficit - 1
        v = inputlength(outlen, self.ratio, self.phi_idx if self.kind != "decimator" else 1, round_up)
        return v + self.input_deficit - 1

    def filt(self, x):
        """filt(self::FIRFilter, x), src/Filters/stream_filt.jl:627-637 (+ the filt! loops :409-560)."""
        x = np.asarray(x)
        if x.ndim != 1:
            raise ArgumentError("FIRFilter filters vectors")
        xdt = _gpu_dtype(_promote(x))
        x = np.ascontiguousarray(x, dtype=xdt)
        if self.kind == "arbitrary":
            return self._filt_arbitrary(x, xdt)
        if xdt not in self._plans:
            self._plans[xdt] = _lib.ResamplePlan(xdt, self.h, self.interpolation, self.decimation)
        plan = self._plans[xdt]
        if self.history is None or self.history.dtype != xdt:
            self.history = np.zeros(self.history_len, dtype=xdt)          # history = zeros(historyLen), :175
        xlen = x.size
        I, D = self.interpolation, self.decimation
        if xlen < self.input_deficit:                                      # :484-488
            self.history = self._shiftin(self.history, x)
            self.input_deficit -= xlen
            return np.zeros(0, dtype=plan.out_dtype)
        phi0 = self.phi_idx - 1 if self.kind in ("rational", "interpolator") else 0
        nout = outputlength(xlen - self.input_deficit + 1, self.ratio, phi0 + 1) if self.kind != "standard" else xlen
        xe = np.concatenate([self.history, x])
        n0 = self.history_len + self.input_deficit - 1                     # index in [history; x] of the first output's newest sample
        out = np.empty(nout, dtype=plan.out_dtype)
        plan.exec(xe, xe.size, 1, n0, phi0, out, nout)
        total = phi0 + nout * D                                            # phase recurrence after nout outputs
        self.input_deficit = self.input_deficit + total // I - xlen        # :511
        if self.kind in ("rational", "interpolator"):
            self.phi_idx = total % I + 1
        if self.kind == "interpolator":
            self.input_deficit = 1                                         # :463
        self.history = self._shiftin(self.history, x)                      # :512
        return out

    def _filt_arbitrary(self, x, xdt):
        """filt!(buffer, ::FIRFilter{FIRArbitrary}, x), src/Filters/stream_filt.jl:579-625.  Output j of the call sits at
        total phase acc + j*delta; the loop `while xIdx <= xLen` produces exactly the j with
        inputDeficit + floor((acc + j*delta) / Nphi) <= xLen, counted here in exact rational arithmetic."""
        if xdt not in self._plans:
            self._plans[xdt] = _lib.ResampleArbPlan(xdt, self.h, self.nphases)
        plan = self._plans[xdt]
        if self.history is None or self.history.dtype != xdt:
            self.history = np.zeros(self.history_len, dtype=xdt)
        xlen = x.size
        if xlen < self.input_deficit:                                                         # :590-594
            self.history = self._shiftin(self.history, x)
            self.input_deficit -= xlen
            return np.zeros(0, dtype=plan.out_dtype)
        nout, new_deficit, new_acc = _arb_advance(self.phi_accumulator, self.input_deficit, self.delta, self.nphases, xlen)
        xe = np.concatenate([self.history, x])
        n0 = self.history_len + self.input_deficit - 1
        out = np.empty(nout, dtype=plan.out_dtype)
        plan.exec(xe, xe.size, n0, self.phi_accumulator, self.delta, out, nout)
        self.input_deficit, self.phi_accumulator = new_deficit, new_acc
        self.phi_idx = 1 + math.floor(self.phi_accumulator)
        self.history = self._shiftin(self.history, x)            # :621
        return out

    @staticmethod
    def _shiftin(a, b):
        """shiftin!, src/util.jl:299-314."""
        if a.size == 0:
            return a
        return np.concatenate([a, b.astype(a.dtype, copy=False)])[-a.size:]


def _arb_advance(acc, input_deficit, delta, nphases, xlen):
    """Bookkeeping of one filt! call of FIRFilter{FIRArbitrary} with xlen >= inputDeficit samples
    (src/Filters/stream_filt.jl:567-625) in exact rational arithmetic: output j sits at total phase acc + j*delta and
    the loop `while xIdx <= xLen` keeps every j with inputDeficit + floor((acc + j*delta) / Nphi) <= xLen.
    Returns (number of outputs, inputDeficit after the call (:620), phiAccumulator after the call)."""
    A, Dl, N = Fraction(acc), Fraction(delta), nphases
    M = xlen - input_deficit + 1
    nout = math.ceil((M * N - A) / Dl)                           # number of j >= 0 with A + j*Dl < M*N
    P = A + nout * Dl                                            # phase after the last update! (:567-577)
    q = P // N
    new_acc = float(P - q * N)
    if new_acc >= N:                                             # rounding of a value just below Nphi
        new_acc = math.nextafter(float(N), 0.0)
    return nout, input_deficit + int(q) - xlen, new_acc


def filt_multirate(h, x, ratio, nphases=32):
    """filt(h::Vector, x::AbstractVector, ratio::Union{Integer,Rational}) and filt(h, x, rate::AbstractFloat, Nphi=32),
    src/Filters/stream_filt.jl:663-672."""
    return FIRFilter(h, ratio, nphases).filt(x)


def _resample_arbitrary(x, rate, h, nphases, dims):
    """resample(x, rate::AbstractFloat[, h, Nphi]; dims), src/Filters/stream_filt.jl:692-704, 751-775: a fresh FIRArbitrary
    filter per column, undelay!, zero-padding to inputlength(outLen, RoundUp), first ceil(length * rate) outputs."""
    if isinstance(x, DeviceArray):
        raise NotImplementedError("arbitrary-rate resampling of a DeviceArray: copy to the host first")
    x = np.asarray(x)
    if not rate > 0.0:
        raise DomainError("rate must be greater than 0")
    sf = FIRFilter(h, rate, nphases)
    if x.ndim > 1:
        if dims is None:
            raise ArgumentError("resample of an array needs `dims`")
        xm = np.moveaxis(x, dims, 0)
    else:
        xm = x
    nx = xm.shape[0]
    cols = xm.reshape(nx, -1)
    outlen = math.ceil(nx * rate)                                                             # :698
    res = None
    for c in range(cols.shape[1]):
        sf.reset()
        sf.setphase(sf.timedelay())                                                           # undelay!, :706-714
        # one sample more than inputlength(outLen, RoundUp) (:699): the extra zero only guarantees that the exact count
        # of outputs reaches outLen; the retained outputs never see it
        npad = max(sf.inputlength(outlen, round_up=True), 0) + 1
        xp = np.zeros(npad, dtype=cols.dtype)
        m = min(nx, npad)
        xp[:m] = cols[:m, c]
        y = sf.filt(xp)
        if y.size < outlen:
            raise AssertionError("Resample output shorter than expected.")                   # :722
        if res is None:
            res = np.empty((outlen, cols.shape[1]), dtype=y.dtype, order="F")
        res[:, c] = y[:outlen]
    if res is None:
        res = np.empty((outlen, 0), dtype=np.float64)
    if x.ndim > 1:
        return np.moveaxis(res.reshape((outlen,) + xm.shape[1:]), 0, dims)
    return res.reshape(outlen)


def resample(x, rate, h=None, nphases=32, dims=None):
    """resample(x, rate[, h]; dims), src/Filters/stream_filt.jl:688-775: Integer / Rational rates through the rational
    polyphase kernel, floating-point rates through FIRArbitrary with `nphases` phases (default 32).
    Output eltype promote_type(eltype(h), eltype(x)) (:654); length ceil(length(x) * rate) (:698)."""
    if isinstance(rate, (float, np.floating)):
        return _resample_arbitrary(x, float(rate), h, int(nphases), dims)
    dev = isinstance(x, DeviceArray)
    if not dev:
        x = np.asarray(x)
    rate = _as_ratio(rate)
    if rate <= 0:
        raise ArgumentError("rate must be positive")
    if h is None:
        h = resample_filter(rate)
    h = np.asarray(h)
    if h.ndim != 1 or h.size == 0:
        raise ArgumentError("h must be a non-empty vector")
    if np.iscomplexobj(h):
        raise NotImplementedError("complex resampling taps are outside the B200 hot-path scope")
    hT = np.ascontiguousarray(h, dtype=np.float32 if h.dtype == np.float32 else np.float64)
    if dev:                                              # device pipeline form (vector)
        if x.ndim != 1:
            raise ArgumentError("device resample takes a vector")
        nout = math.ceil(x.shape[0] * rate)
        n0, phi0 = resample_phase(hT.size, rate)
        plan = _lib.ResamplePlan(x.dtype, hT, rate.numerator, rate.denominator)
        out = DeviceArray((nout,), plan.out_dtype)
        plan.exec_dev(x.ptr, x.shape[0], 1, n0, phi0, out.ptr, nout, 0)
        from .device import sync
        sync()
        plan.close()
        return out
    if x.ndim > 1:
        if dims is None:
            raise ArgumentError("resample of an array needs `dims`")
        xm = np.moveaxis(x, dims, 0)
    else:
        xm = x
    xdt = _gpu_dtype(_promote(xm))
    xF, nx, ncols = _cols(xm, xdt)
    nout = math.ceil(nx * rate)
    n0, phi0 = resample_phase(hT.size, rate)
    plan = _lib.ResamplePlan(xdt, hT, rate.numerator, rate.denominator)
    res = np.empty((nout, ncols), dtype=plan.out_dtype, order="F")
    plan.exec(xF, nx, ncols, n0, phi0, res, nout)
    plan.close()
    if x.ndim > 1:
        res = res.reshape((nout,) + xm.shape[1:])
        return np.moveaxis(res, 0, dims)
    return res.reshape(nout)
