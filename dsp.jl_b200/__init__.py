"""dspb200 -- B200-native (sm_100a) implementation of DSP.jl's data-parallel hot path.

Host-side mirror of the reference's call signatures for that path (the Julia glue in julia/DSPB200.jl is
the same thin layer over the same C ABI, include/dspb200.h):

    filt, filt_, conv, conv_, optimalfftfiltlength                    (src/dspbase.jl)
    fftfilt, fftfilt_, tdfilt, tdfilt_, resample, resample_filter      (src/Filters)
    periodogram, welch_pgram, welch_pgram_, WelchConfig, spectrogram, stft, power, freq, time
                                                                       (src/periodograms.jl)
    hanning, hamming, rect, bartlett, kaiser, nextfastfft              (src/windows.jl, src/util.jl)

All numerics run in hand-written CUDA kernels inside libdspb200.so; there is no CPU fallback.
The directory is named `dsp.jl_b200` (not importable as written), so it is loaded under the module
name `dspb200` by the shim `dspb200.py` at the repository root.
"""
from . import _lib
from ._lib import DSPB200Error, device_count, launch_count
from .device import DeviceArray, sync, to_device, to_host
from .errors import ArgumentError, DimensionMismatch, DomainError
from .util import fftabs2type, fftintype, fftouttype, nextfastfft, rfftfreq, fftfreq
from .windows import bartlett, hamming, hann, hanning, kaiser, rect
from .dspbase import SMALL_FILT_CUTOFF, conv, conv_, filt, filt_, optimalfftfiltlength, os_fft_complexity
from .filters import (FIRFilter, fftfilt, fftfilt_, filt_multirate, inputlength, kaiserord, outputlength, resample,
                      resample_filter, resample_phase, tdfilt, tdfilt_)
from .filters import filt_ as filt_hx_
from .periodograms import (Periodogram, Periodogram2, Spectrogram, WelchConfig, arraysplit, arraysplit_count, compute_window, fftshift,
                           filt_welch, freq, periodogram, power, spectrogram, stft, time, welch_pgram, welch_pgram_)

from .multitaper import (Coherence, CrossPowerSpectra, MTConfig, MTCrossSpectraConfig, dpss, dpss_config, dpsseig,
                         mt_coherence, mt_cross_power_spectra, mt_pgram, mt_spectrogram)
from .clients import alignsignals, filtfilt, finddelay, hilbert, shiftsignal, xcorr
from . import device, filters, sharding

__version__ = "0.1.0"
