"""CPU oracle for the DSP.jl hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / scipy-pocketfft) of the reference's
algorithm for the path named by BASELINE.json `north_star`: FIR filtering,
overlap-save convolution, Welch / spectrogram / STFT / periodogram estimation
and rational polyphase resampling.  Every function cites the reference
file:line (relative to /root/reference, DSP.jl v0.8.5 @ 2d57c27) it follows.

Rules:
  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline`
    / `--impl reference` legs may import this package -- and only as the
    checker / the timed CPU baseline, never as the product.  The product
    package (`dsp.jl_b200/`, imported as `dspb200`) must never import it and
    has no CPU fallback.
  * the reference itself (Julia + FFTW.jl) cannot run in the build container
    (no `julia`, no libfftw3).  The FFT arithmetic lives in the third-party,
    un-vendored FFTW.jl (Project.toml compat "1.8", libfftw3 3.3.x, no
    Manifest committed); it is restated with scipy.fft (pocketfft), which is a
    correct DFT to ~1 ulp -- the reference's own tests never pin FFTW rounding.
  * PINNING: the oracle is checked in tests/test_oracle_golden.py against every
    golden vector / known-answer test the reference holds for this path
    (tests/golden/reference_goldens.npz is produced from
    /root/reference/test/data by tests/golden/import_reference_goldens.py;
    inline MATLAB KATs are restated in the test file with their file:line).

Precision modes: functions compute in the dtype implied by the reference's
promotion rules (Float32 in -> Float32 arithmetic, incl. a Float32 FFT) unless
`f64=True` is passed, which evaluates the same formula in double precision and
is used as ground truth for the norm-relative error budget.
"""
from . import util, windows, dspbase, filters, periodograms  # noqa: F401
