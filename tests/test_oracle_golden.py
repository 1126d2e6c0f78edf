"""Pin the CPU oracle against the reference's golden vectors and known-answer tests.

Every case cites the reference test (file:line under /root/reference) it restates.  `≈` is Julia's
isapprox (norm-relative, rtol = sqrt(eps))."""
import numpy as np
import pytest

from conftest import approx, relerr
from oracle import dspbase as od
from oracle import filters as of
from oracle import periodograms as op
from oracle import util as ou
from oracle import windows as ow

RNG = np.random.default_rng(1776)


# ------------------------------------------------------------------ windows / util

def test_windows_matlab_goldens(goldens):
    # test/windows.jl:50-73
    assert np.array_equal(ow.rect(128), np.ones(128))
    assert approx(ow.hanning(128), goldens["hanning128"])
    assert approx(ow.hamming(128), goldens["hamming128"])
    assert approx(ow.bartlett(128), goldens["bartlett128"])
    assert approx(ow.kaiser(128, 0.4 / np.pi), goldens["kaiser128_0.4"])   # test/windows.jl:102-104
    assert ow.hanning(1)[0] == 1.0
    assert ow.hanning(8)[0] == 0.0 and ow.hanning(8)[-1] == 0.0


def test_nextfastfft():
    # test/util.jl:55-60
    assert ou.nextfastfft(64) == 64
    assert ou.nextfastfft(65) == 70
    assert ou.nextfastfft(127) == 128
    assert [ou.nextfastfft(n) for n in (1, 2, 3, 11, 13, 1000, 1001)] == [1, 2, 3, 12, 14, 1000, 1008]


def test_optimalfftfiltlength():
    # test/dsp.jl:39 and SURVEY.md A4 values
    assert od.optimalfftfiltlength(1, 3) == 1
    assert od.optimalfftfiltlength(257, 2 ** 20) == 2048
    assert od.optimalfftfiltlength(4097, 2 ** 26) == 65536
    assert od.optimalfftfiltlength(67, 2 ** 20) == 512
    assert od.optimalfftfiltlength(127, 2 ** 18 - 1) == 1024


# ------------------------------------------------------------------ filt

def test_filt_exact_small():
    # test/dsp.jl:10-21
    b = np.array([1., 2., 3., 4.])
    x = np.array([1., 1., 0., 1., 1., 0., 0., 0.])
    assert np.array_equal(od.filt(b, 1., x), [1., 3., 5., 8., 7., 5., 7., 4.])
    assert np.array_equal(od.filt(b, 1., np.arange(1.0, 11.0)), [1., 4., 10., 20., 30., 40., 50., 60., 70., 80.])
    x2 = np.stack([x, np.arange(1.0, 9.0)], axis=1)
    y2 = od.filt(b, 1., x2)
    assert np.array_equal(y2[:, 0], od.filt(b, 1., x)) and np.array_equal(y2[:, 1], od.filt(b, 1., np.arange(1.0, 9.0)))
    with pytest.raises(ValueError):
        od.filt(np.zeros(0), 1., x)
    with pytest.raises(ValueError):
        od.filt(b, 0., x)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_filt_vectorised_matches_literal_loop(dt):
    # the vectorised nested-fma form is the literal transposed-direct-form loop (src/dspbase.jl:95-105)
    b = RNG.standard_normal(9).astype(dt)
    x = RNG.standard_normal(64).astype(dt)
    lit = od.filt_fir_literal(b, x)
    vec = od.filt(b, np.ones(1, dtype=dt), x)
    assert vec.dtype == dt
    if dt is np.float32:
        assert np.array_equal(lit, vec)
    else:
        assert relerr(lit, vec) < 4e-16


def test_fftfilt_filt_tdfilt_agree():
    # test/filt.jl:312-331: xlen in 2^(7:18)-1, blen in 2^(1:7)-1; fftfilt ≈ filt(b,[1.0],x) ≈ filt(b,x)
    for xlen in (2 ** 7 - 1, 2 ** 10 - 1, 2 ** 13 - 1):
        for blen in (2 ** 1 - 1, 2 ** 4 - 1, 2 ** 7 - 1):
            b = RNG.standard_normal(blen)
            for x in (RNG.standard_normal(xlen), RNG.standard_normal((xlen, 2))):
                ref = od.filt(b, [1.0], x)
                assert approx(of.fftfilt(b, x), ref)
                assert approx(of.filt(b, x), ref)
                assert approx(of.tdfilt(b, x), ref)


# ------------------------------------------------------------------ conv

def test_conv_exact_integers():
    # test/dsp.jl:53-59, 80-81
    a = np.array([1, 2, 1, 2])
    b = np.array([1, 2, 3])
    exp = [1, 4, 8, 10, 7, 6]
    assert np.array_equal(od.conv(a, b), exp)
    assert np.array_equal(od.conv(a.astype(np.float64), b.astype(np.float64)), exp)
    assert np.array_equal(od.conv(a * 1j, b.astype(np.complex128)).imag, exp)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_conv_algorithms_agree(dt):
    # test/dsp.jl:98-121: M,N in {10,200}; all algorithms ≈
    for M in (10, 200):
        for N in (10, 200):
            u = RNG.standard_normal(M).astype(dt)
            v = RNG.standard_normal(N).astype(dt)
            if dt is np.complex128:
                u = u + 1j * RNG.standard_normal(M)
                v = v + 1j * RNG.standard_normal(N)
            ref = od.conv(u, v, "direct")
            for alg in ("fft_simple", "fft_overlapsave", "fft", "fast", "auto"):
                assert approx(od.conv(u, v, alg), ref), (M, N, alg)
    with pytest.raises(ValueError):
        od.conv(np.ones(300), np.ones(300), "bogus")


def test_conv_empty():
    # test/dsp.jl:41-49
    for alg in ("direct", "fft", "fft_simple", "fft_overlapsave"):
        assert np.array_equal(od.conv(RNG.standard_normal(5), np.zeros(0), alg), np.zeros(4))
        assert od.conv(np.zeros(0), np.zeros(0), alg).size == 0


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex128])
def test_os_kernel_vs_single_fft(dt):
    # test/dsp.jl:271-314 (N=1): regular and adversarial (nsmall, nfft), "three padded blocks" case
    cases = [(128, 12, od.optimalfftfiltlength(12, 128)), (128, 128, od.optimalfftfiltlength(128, 128)),
             (128, 12, 256), (128, 13, 32), (128, 12, 32), (25, 4, 16)]
    for nu, nv, nfft in cases:
        u = RNG.standard_normal(nu).astype(dt)
        v = RNG.standard_normal(nv).astype(dt)
        if np.issubdtype(dt, np.complexfloating):
            u = u + 1j * RNG.standard_normal(nu)
            v = v + 1j * RNG.standard_normal(nv)
        os_out = od.conv_kern_os(u, v, nfft)
        assert os_out.dtype == dt
        assert approx(os_out, od.conv_kern_fft(u, v)), (nu, nv, nfft)
        assert relerr(os_out, od.conv_exact(u, v)) < (2e-6 if dt is np.float32 else 1e-13)


# ------------------------------------------------------------------ periodograms

DATA = np.arange(8)
DATA0 = np.array([98.0, 13.656854249492380, 4.0, 2.343145750507620, 2.0, 2.343145750507620, 4.0, 13.656854249492380])


def test_periodogram_welch_spectrogram_0to7():
    # test/periodograms.jl:92-106
    assert approx(op.periodogram(DATA, onesided=False)[0], DATA0)
    assert approx(op.welch_pgram(DATA, 8, 0, onesided=False)[0], DATA0)
    assert approx(op.spectrogram(DATA, 8, 0, onesided=False)[0][:, 0], DATA0)
    z = DATA + 1j * DATA
    assert approx(op.periodogram(z, onesided=False)[0], DATA0 * 2)
    assert approx(op.welch_pgram(z, 8, 0, onesided=False)[0], DATA0 * 2)
    assert approx(op.spectrogram(z, 8, 0, onesided=False)[0][:, 0], DATA0 * 2)


@pytest.mark.parametrize("n,nov,expected", [(2, 0, [34.5, 0.5]), (3, 0, [25.5, 1.0, 1.0]),
                                            (3, 1, [35.0, 1.0, 1.0]), (4, 1, [45, 2, 1, 2])])
def test_welch_rect_kats(n, nov, expected):
    # test/periodograms.jl:108-131 (MATLAB pwelch)
    assert approx(op.welch_pgram(DATA, n, nov, onesided=False)[0], np.array(expected, float))
    assert approx(op.spectrogram(DATA, n, nov, onesided=False)[0].mean(axis=1), np.array(expected, float))


def test_windowed_periodogram_kats():
    # test/periodograms.jl:139-166
    cases = ((ow.hamming, [65.461623986801527, 20.556791795515764, 0.369313143650544, 0.022167446610882,
                           0.025502985564107, 0.022167446610882, 0.369313143650544, 20.556791795515764]),
             (ow.bartlett, [62.999999999999993, 21.981076052592442, 0.285714285714286, 0.161781090264695,
                            0.142857142857143, 0.161781090264695, 0.285714285714286, 21.981076052592442]))
    for win, exp in cases:
        exp = np.array(exp)
        for w in (win, win(8)):
            assert approx(op.periodogram(DATA, window=w, onesided=False)[0], exp, rtol=1e-8)
            assert approx(op.welch_pgram(DATA, 8, 0, window=w, onesided=False)[0], exp, rtol=1e-8)
            assert approx(op.spectrogram(DATA, 8, 0, window=w, onesided=False)[0][:, 0], exp, rtol=1e-8)


def test_padded_periodogram_kats():
    # test/periodograms.jl:168-222
    exp = np.array([98, 174.463067389405, 121.968086934209, 65.4971744936088, 27.3137084989848, 12.1737815028909,
                    10.3755170959439, 10.4034038628775, 8, 5.25810953219633, 4.47015397150535, 4.89522578856669,
                    4.68629150101524, 3.69370284475603, 3.1862419983415, 3.61553458569862, 2])
    assert approx(op.periodogram(DATA, nfft=32)[0], exp)
    assert approx(op.welch_pgram(DATA, 8, 0, nfft=32)[0], exp)
    assert approx(op.spectrogram(DATA, 8, 0, nfft=32)[0][:, 0], exp)
    exph = np.array([65.4616239868015, 122.101693164395, 98.8444689598445, 69.020252632913, 41.1135835910315,
                     20.5496474310966, 8.43291449161938, 2.78001620362588, 0.738626287301088, 0.174995741770789,
                     0.0501563022944516, 0.0327357460012861, 0.0443348932217643, 0.0553999745503552,
                     0.0561319901616643, 0.0526025934871384, 0.0255029855641069])
    assert approx(op.periodogram(DATA, window=ow.hamming, nfft=32)[0], exph)
    assert approx(op.welch_pgram(DATA, 8, 0, window=ow.hamming, nfft=32)[0], exph)
    assert approx(op.spectrogram(DATA, 8, 0, window=ow.hamming, nfft=32)[0][:, 0], exph)


def test_spectrogram_matlab_golden(goldens):
    # test/periodograms.jl:25-36
    p, f, t = op.spectrogram(goldens["spectrogram_x"], 256, 128, fs=10)
    assert approx(p, goldens["spectrogram_p"])
    assert approx(f, goldens["spectrogram_f"])
    assert approx(t, goldens["spectrogram_t"])
    assert relerr(p, goldens["spectrogram_p"]) < 1e-13


def test_stft_matlab_golden(goldens):
    # test/periodograms.jl:332-344
    S = op.stft(goldens["stft_x"], 400, 400 - 160, nfft=512, fs=16000, window=ow.hanning)
    Sml = goldens["stft_S_real"] + 1j * goldens["stft_S_imag"]
    assert S.shape == (257, 29)
    assert approx(S, Sml)
    assert np.max(np.abs(S - Sml)) < 1e-12


def test_fft2oneortwosided():
    # test/periodograms.jl:346-379
    import scipy.fft as sfft
    n = 10
    for nfft in (n, n + 2, n + 3):
        x = np.zeros(nfft)
        x[:n] = RNG.random(n)
        two = op.fft2oneortwosided(sfft.rfft(x)[None, :], nfft, False)[0]
        assert approx(two, sfft.fft(x))
        one = op.fft2oneortwosided(sfft.rfft(x)[None, :], nfft, True)[0]
        assert approx(one, sfft.rfft(x))


def test_arraysplit():
    # test/periodograms.jl:393-396 (#124) and docstring examples src/periodograms.jl:96-113
    q = op.arraysplit(np.ones(1000), 100, 10)
    assert q.shape[0] == 11 and np.array_equal(q.mean(axis=1), np.ones(11))
    a = op.arraysplit(np.array([0.1, 0.2, 0.3, 0.4, 0.5]), 3, 2, 8)
    assert a.shape == (3, 8) and np.array_equal(a[2, :3], [0.3, 0.4, 0.5]) and not a[:, 3:].any()
    b = op.arraysplit(np.array([0.1, 0.2, 0.3, 0.4, 0.5]), 3, 1, 3, np.array([1, 2, 1]))
    assert np.allclose(b, [[0.1, 0.4, 0.3], [0.3, 0.8, 0.5]])
    with pytest.raises(op.DomainError):
        op.arraysplit(np.ones(10), 4, 4)
    with pytest.raises(op.DomainError):
        op.arraysplit(np.ones(10), 4, 2, 3)


def test_welch_float32_sequential_vs_pairwise():
    # hard part 2: the reference's sequential Float32 accumulation is itself ~1e-6 off the f64 truth
    x = RNG.standard_normal(1 << 16).astype(np.float32)
    truth = op.welch_pgram(x, 256, 128, window=ow.hanning, f64=True)[0]
    seq = op.welch_pgram(x, 256, 128, window=ow.hanning, sequential=True)[0]
    par = op.welch_pgram(x, 256, 128, window=ow.hanning)[0]
    assert seq.dtype == np.float32 and par.dtype == np.float32
    assert relerr(seq, truth) < 5e-6
    assert relerr(par, truth) < 1e-6


# ------------------------------------------------------------------ resample

@pytest.mark.parametrize("rate", ["1/2", "2/1", "3/2", "2/3"])
def test_resample_matlab_goldens(goldens, rate):
    # test/resample.jl:8-24
    from fractions import Fraction
    r = Fraction(rate)
    key = f"{r.numerator}_{r.denominator}"
    x, h, y = goldens["resample_x"], goldens[f"resample_taps_{key}"], goldens[f"resample_y_{key}"]
    ylit = of.resample_literal(x, r, h)
    assert approx(ylit, y)
    yvec = of.resample(x, r, h)
    assert yvec.shape == y.shape and approx(yvec, y)
    assert relerr(yvec, ylit) < 1e-15
    assert approx(of.resample(x, r), y, rtol=1e-3)      # default taps ok


def test_resample_exact_tiny():
    # test/filt_stream.jl:366-367 (round-half-even in setphase!)
    h = np.array([0, 0, 1, 0, 0, 0.])
    assert np.array_equal(of.resample_literal(np.array([1., 2.]), 3, h), [1, 0, 0, 2, 0, 0])
    assert np.array_equal(of.resample_literal(np.array([1., 2.]), "3/2", h), [1, 0, 0])
    assert np.array_equal(of.resample(np.array([1., 2.]), 3, h), [1, 0, 0, 2, 0, 0])
    assert np.array_equal(of.resample(np.array([1., 2.]), "3/2", h), [1, 0, 0])


def test_resample_closed_form_matches_stateful_loops():
    # test/filt_stream.jl:231-281 grid (subset): the closed form equals the stateful reference loops
    from fractions import Fraction
    for interp in (1, 5, 14):
        for dec in (1, 9, 17):
            if interp == dec:
                continue
            r = Fraction(interp, dec)
            h = RNG.standard_normal(56)
            x = RNG.standard_normal(301) + 1j * RNG.standard_normal(301)
            a = of.resample_literal(x, r, h)
            b = of.resample(x, r, h)
            assert a.shape == b.shape, (interp, dec)
            assert relerr(a, b) < 1e-14, (interp, dec)


def test_default_resample_taps():
    # SURVEY.md A9 / BASELINE.md: 3//2 -> 111 taps (37 per phase), DC gain = Nphi
    h = of.resample_filter("3/2")
    assert len(h) == 111 and abs(h.sum() - 3.0) < 1e-12 and np.allclose(h, h[::-1])
    n, alpha = of.kaiserord(0.2 / 3)
    assert n == 110 and abs(alpha - 1.7995) < 1e-3


def test_resample_polyphase_vs_naive():
    # test/filt_stream.jl:3-17: zero-stuff + filt + downsample
    from fractions import Fraction
    h = RNG.standard_normal(23)
    x = RNG.standard_normal(97)
    for r in (Fraction(5, 9), Fraction(14, 9), Fraction(5, 1)):
        I, D = r.numerator, r.denominator
        sf = of.FIRFilterState(h, r)
        y = sf.filt(x)
        up = np.zeros(len(x) * I)
        up[::I] = x
        naive = od.filt(h, 1.0, up)[::D]
        assert approx(y, naive[:len(y)]) and abs(len(y) - len(naive)) <= 1


def test_multitaper_matlab_goldens(goldens):
    # test/windows.jl:32-36, test/periodograms.jl:381-386, 404-470 (MATLAB dpss / pmtm)
    assert approx(ow.dpss(128, 4), goldens["dpss128_4"])
    assert approx(op.mt_pgram(goldens["stft_x"], fs=16000)[0], goldens["mt_pgram"])
    x = goldens["pmtm_x"]
    nfft = 1 << (x.size - 1).bit_length()
    p, f = op.mt_pgram(x, fs=1000, nw=4, nfft=nfft)
    assert approx(p, goldens["pmtm_pxx"]) and approx(f, goldens["pmtm_fx"])
    z = x + 1j * goldens["pmtm_y"]
    p, f = op.mt_pgram(z, fs=1000, nw=4, nfft=nfft)
    m = (f > 0) & (f < 500)
    assert approx(p[m], goldens["pmtm_pzz"][1:m.sum() + 1])
    x0 = goldens["spectrogram_x"]
    mt, fq, tm = op.mt_spectrogram(x0, 256, 128, fs=10)
    assert approx(mt[:, 0], op.mt_pgram(x0[:256], fs=10)[0]) and approx(tm, goldens["spectrogram_t"])


def test_hilbert_reference_properties():
    # test/util.jl:4-50 pins hilbert() through exact analytic properties
    from oracle.util import hilbert
    t = np.arange(0, 2, 1 / 256)
    a0, a1, a2, a3 = np.sin(np.pi * t), np.cos(np.pi * t), np.sin(2 * np.pi * t), np.cos(2 * np.pi * t)
    a = np.stack([a0, a1, a2, a3], axis=1)
    h = hilbert(a)
    assert np.allclose(h.real, a) and np.allclose(np.abs(h), 1.0)
    assert np.allclose(np.angle(h[:256, 0]), -np.pi / 2 + np.pi / 256 * np.arange(256))
    assert np.allclose(np.angle(h[:256, 1]), np.pi / 256 * np.arange(256))
    assert np.allclose(np.angle(h[:128, 2]), -np.pi / 2 + np.pi / 128 * np.arange(128))
    assert np.allclose(np.angle(h[:128, 3]), np.pi / 128 * np.arange(128))
    assert np.allclose(h[:, 1].imag, a0)
    odd = np.concatenate([np.ones(10), np.zeros(9)])
    assert np.allclose(hilbert(odd).real, odd)
    r = np.random.default_rng(3).integers(1, 21, 128)
    assert np.array_equal(hilbert(r), hilbert(r.astype(np.float64)))
    assert hilbert(a0.astype(np.float32)).dtype == np.complex64


def _mne_case(goldens):
    fs, n = 1000.0, 1024
    t = np.arange(n) / fs
    sin_1 = np.sin(2 * np.pi * 12.0 * t)
    sin_2 = np.sin(np.pi * (2 * 12.0 * t + 1))
    return fs, n, sin_1, sin_2


def test_mt_cross_power_spectra_mne_golden(goldens):
    # test/multitaper.jl:277-300: MNE-python csd_array_multitaper, low_bias tapers weighted by eigenvalue, demeaned
    from oracle import periodograms as op
    fs, n, sin_1, sin_2 = _mne_case(goldens)
    win, w = op.dpss_config(n, keep_only_large_evals=True, weight_by_evals=True)
    cs, f = op.mt_cross_power_spectra(np.stack([sin_1, sin_2]), fs=fs, window=win, taper_weights=w, demean=True)
    ref = (goldens["csd_mt_values_re"] + 1j * goldens["csd_mt_values_im"]).reshape((512, 2, 2)).transpose(2, 1, 0)
    assert np.allclose(f[1:], goldens["csd_mt_frequencies"])
    assert np.allclose(cs[:, :, 1:], ref, rtol=1.5e-8, atol=0)


def test_mt_coherence_mne_kat(goldens):
    # test/multitaper.jl:254-275: MNE-python spectral_connectivity(method="coh", fmin=10, fmax=15)
    from oracle import periodograms as op
    fs, n, sin_1, _ = _mne_case(goldens)
    win, w = op.dpss_config(n, keep_only_large_evals=True, weight_by_evals=True)
    sig = np.stack([sin_1, sin_1 + 3 * goldens["mt_noise"]])
    coh, f = op.mt_coherence(sig, fs=fs, window=win, taper_weights=w, demean=True, freq_range=(10, 15))
    assert np.all((f >= 10) & (f <= 15))
    assert abs(coh.mean(axis=2)[1, 0] - 0.982356762670818) < 1e-12
    assert np.array_equal(coh, np.transpose(coh, (1, 0, 2))) and np.all(coh[0, 0] == 1) and np.all(coh[1, 1] == 1)


def _naive_arbitrary(h, x, rate, nphi):
    """naivefilt(h, x, resamplerate::AbstractFloat, numfilters), test/filt_stream.jl:19-43: interpolate by nphi with the
    plain polyphase filter, then pick samples with linear interpolation."""
    from fractions import Fraction
    from oracle import filters as of
    xi = of.FIRFilterState(h, Fraction(nphi, 1)).filt(x)
    y, xidx, alpha = [], 1, 0.0
    delta, stride = np.modf(nphi / rate)
    stride = int(stride)
    while xidx < len(xi):
        lo, up = xi[xidx - 1], xi[xidx]
        y.append(lo + alpha * (up - lo))
        alpha += delta
        xidx += int(np.floor(alpha)) + stride
        alpha = alpha % 1.0
    return np.asarray(y)


@pytest.mark.parametrize("rate", [0.7312, 1.2957, 2.618])
def test_arbitrary_resampler_literal_vs_naive(rate):
    # test/filt_stream.jl:288-330: stateless == stateful == piecewise ~ naive (approx: rtol sqrt(eps))
    from oracle import filters as of
    # (the reference test uses a long, narrow-transition filter: FIRArbitrary's derivative-bank step at the last phase drops
    #  the h[1]*x[next] term, so the comparison is only that tight when the edge taps are negligible)
    nphi = 32
    h = of.resample_filter_arb(rate, nphi, 1.0, 200)
    x = np.random.default_rng(7).standard_normal(101)
    naive = _naive_arbitrary(h, x, rate, nphi)
    stateless = of.FIRArbitraryState(h, rate, nphi).filt(x)
    sf = of.FIRArbitraryState(h, rate, nphi)
    piecewise = np.concatenate([sf.filt(x[i:i + 1]) for i in range(len(x))])
    n = min(len(naive), len(stateless), len(piecewise))
    assert n >= int(len(x) * rate) - 2
    assert np.allclose(naive[:n], stateless[:n], rtol=1.5e-8, atol=1e-10)
    assert np.array_equal(stateless[:n], piecewise[:n])


def test_arbitrary_resample_lengths():
    # test/resample.jl:99-107 (issue 317): buffer length bookkeeping of resample(x, rate::Float64)
    from oracle import filters as of
    assert len(of.resample_arb_literal(np.sin(np.arange(1.0, 35547.0)), 1 / 55.55)) == 640
    assert len(of.resample_arb_literal(np.random.default_rng(0).standard_normal(1822), 0.9802414928649835)) == 1786
    assert np.array_equal(of.resample_arb_literal(np.zeros(1000), 0.012), np.zeros(12))


def test_conv_nd_oracle_known_answers():
    # test/dsp.jl:130-165, 225-252: the reference's 2-D and 3-D integer known answers pin the N-D restatements
    from oracle import dspbase as od
    a = np.array([[1, 2, 1], [2, 3, 1], [1, 2, 1]])
    b = np.array([[3, 2], [0, 1]])
    expectation = np.array([[3, 8, 7, 2], [6, 14, 11, 3], [3, 10, 10, 3], [0, 1, 2, 1]])
    assert np.array_equal(od.conv_td_nd(a, b), expectation)
    assert np.allclose(od.conv_kern_fft_nd(a.astype(float), b.astype(float)), expectation, atol=1e-13)
    im = np.array([[3, 5, 5, 2], [3, 6, 6, 3], [3, 6, 6, 3], [0, 1, 1, 1]])
    assert np.allclose(od.conv_kern_fft_nd(a + 1j, b + 0j), expectation + 1j * im, atol=1e-13)
    a3 = np.arange(1, 28).reshape((3, 3, 3), order="F")
    exp3 = np.array([1, 3, 5, 3, 5, 12, 16, 9, 11, 24, 28, 15, 7, 15, 17, 9, 11, 24, 28, 15, 28, 60, 68, 36, 40, 84, 92, 48, 23, 48, 52,
                     27, 29, 60, 64, 33, 64, 132, 140, 72, 76, 156, 164, 84, 41, 84, 88, 45, 19, 39, 41, 21, 41, 84, 88, 45, 47, 96, 100,
                     51, 25, 51, 53, 27]).reshape((4, 4, 4), order="F")
    assert np.array_equal(od.conv_td_nd(a3, np.ones((2, 2, 2), dtype=np.int64)), exp3)
    assert np.allclose(od.conv_kern_fft_nd(a3.astype(float), np.ones((2, 2, 2))), exp3, atol=1e-12)


def test_conv_nd_overlap_save_oracle():
    # test/dsp.jl:270-313 ("Overlap-Save"): unsafe_conv_kern_os! == _conv_kern_fft! for N = 1, 2, 3 with
    # nffts = optimalfftfiltlength(nsmall, nlarge) per dimension, plus the adversarial (nsmall, nfft) pairs; the N-D
    # restatement must also reproduce the 1-D one bit for bit and the reference's integer known answers
    from oracle import dspbase as od
    rng = np.random.default_rng(5)
    for nd, nlarge in ((1, 128), (2, 128), (3, 32)):
        for dt in (np.float32, np.float64, np.complex128):
            for nsmall in (12, nlarge):
                nfft = od.optimalfftfiltlength(nsmall, nlarge)
                u = rng.standard_normal((nlarge,) * nd).astype(dt)
                v = rng.standard_normal((nsmall,) * nd).astype(dt)
                a = od.conv_kern_os_nd(u, v, (nfft,) * nd)
                b = od.conv_kern_fft_nd(u, v)
                assert a.dtype == b.dtype and relerr(a, b) < (2e-6 if dt is np.float32 else 1e-13), (nd, dt, nsmall)
    for nl, ns, nfft in ((128, 12, 256), (128, 13, 32), (128, 12, 32), (25, 4, 16)):
        u, v = rng.standard_normal(nl), rng.standard_normal(ns)
        assert np.array_equal(od.conv_kern_os_nd(u, v, (nfft,)), od.conv_kern_os(u, v, nfft))
        assert relerr(od.conv_kern_os_nd(u, v, (nfft,)), od.conv_exact(u, v)) < 1e-13
    a = np.array([[1, 2, 1], [2, 3, 1], [1, 2, 1]], dtype=float)
    b = np.array([[3, 2], [0, 1]], dtype=float)
    assert np.allclose(od.conv_kern_os_nd(a, b, (2, 4)), [[3, 8, 7, 2], [6, 14, 11, 3], [3, 10, 10, 3], [0, 1, 2, 1]], atol=1e-13)
    u, v = rng.standard_normal((4, 7, 1)), rng.standard_normal((3, 3, 3))            # size(v) > size(u) in one dimension
    assert relerr(od.conv_kern_os_nd(u, v, (8, 8, 4)), od.conv_td_nd(u, v)) < 1e-13


def test_periodogram2_octave_goldens(goldens):
    # test/periodograms.jl:270-330: Octave raPsd2d radial sum / mean, fft2 identity, doc examples, sparse non-square case
    from oracle import periodograms as op
    x = goldens["per2dx"]
    assert np.allclose(op.periodogram2(x, fs=1.0, radialsum=True)[0], goldens["per2dsum"], rtol=1.5e-8)
    assert np.allclose(op.periodogram2(x, fs=1.0, radialavg=True)[0], goldens["per2dmean"], rtol=1.5e-8)
    assert np.allclose(op.periodogram2(x)[0], np.abs(np.fft.fft2(x)) ** 2 / x.size)
    assert np.allclose(op.periodogram2(np.array([[1, 3], [0, 1]]), radialsum=True)[0], [6.25, 4.75])
    assert np.allclose(op.periodogram2(np.array([[1, 3], [0, 1]]), radialavg=True)[0], [6.25, 1.5833333333333333])
    assert np.allclose(op.periodogram2(np.array([[1, 1], [0, 1], [0, 0]]), nfft=(3, 2))[0],
                       [[1.5, 1 / 6], [0.5, 1 / 6], [0.5, 1 / 6]])
