#!/usr/bin/env python
"""Build tests/golden/reference_goldens.npz from the reference's own golden test data.

Run in the build container (where /root/reference exists):
    python tests/golden/import_reference_goldens.py

The text files under /root/reference/test/data are MATLAB/Octave outputs loaded by the reference's
tests through `read_reference_data` (test/FilterTestHelpers.jl:8).  Only the vectors that pin the
hot path are imported; they are stored as float64 arrays, losslessly parsed.

  spectrogram_{x,f,t,p}   test/periodograms.jl:25-36    spectrogram(x, 256, 128; fs=10)
  stft_x, stft_S_{real,imag}  test/periodograms.jl:332-344  stft(x, 400, 240; nfft=512, fs=16000, window=hanning)
  hanning128, hamming128, bartlett128, kaiser128_0.4   test/windows.jl:55-73
  resample_x, resample_taps_I_D, resample_y_I_D   test/resample.jl:8-24
  mt_pgram, pmtm_{x,y,fx,pxx,fz,pzz}   test/periodograms.jl:381-490 (MATLAB pmtm);  dpss128_4   test/windows.jl:30-40
  per2dx, per2dsum, per2dmean   test/periodograms.jl:270-282 (Octave raPsd2d)
  csd_mt_{frequencies,values_re,values_im}, mt_noise   test/multitaper.jl:254-300 (MNE-python csd_array_multitaper / noise for the coherence KAT)
"""
import os
import numpy as np

SRC = "/root/reference/test/data"
HERE = os.path.dirname(os.path.abspath(__file__))

FILES = {
    "spectrogram_x": "spectrogram_x.txt", "spectrogram_f": "spectrogram_f.txt",
    "spectrogram_t": "spectrogram_t.txt", "spectrogram_p": "spectrogram_p.txt",
    "stft_x": "stft_x.txt", "stft_S_real": "stft_S_real.txt", "stft_S_imag": "stft_S_imag.txt",
    "hanning128": "hanning128.txt", "hamming128": "hamming128.txt", "bartlett128": "bartlett128.txt",
    "kaiser128_0.4": "kaiser128,0.4.txt",
    "resample_x": "resample_x.txt",
    # multitaper: test/periodograms.jl:381-490
    "mt_pgram": "mt_pgram.txt", "pmtm_x": "pmtm_x.txt", "pmtm_y": "pmtm_y.txt", "pmtm_fx": "pmtm_fx.txt",
    "pmtm_pxx": "pmtm_pxx.txt", "pmtm_fz": "pmtm_fz.txt", "pmtm_pzz": "pmtm_pzz.txt", "dpss128_4": "dpss128,4.txt",
    # multitaper cross spectra / coherence (MNE-python outputs): test/multitaper.jl:254-300
    "csd_mt_frequencies": "csd_array_multitaper_frequencies.txt", "csd_mt_values_re": "csd_array_multitaper_values_re.txt",
    "csd_mt_values_im": "csd_array_multitaper_values_im.txt", "mt_noise": "noise.txt",
    # 2-D periodogram (Octave raPsd2d): test/periodograms.jl:270-282
    "per2dx": "per2dx.txt", "per2dsum": "per2dsum.txt", "per2dmean": "per2dmean.txt",
}
for r in ("1_2", "2_1", "3_2", "2_3"):
    FILES[f"resample_taps_{r}"] = f"resample_taps_{r}.txt"
    FILES[f"resample_y_{r}"] = f"resample_y_{r}.txt"


def load(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(tok) for tok in line.replace("\t", " ").split()])
    a = np.array(rows, dtype=np.float64)
    return a[:, 0] if a.shape[1] == 1 else (a[0] if a.shape[0] == 1 else a)


if __name__ == "__main__":
    out = {k: load(os.path.join(SRC, v)) for k, v in FILES.items()}
    np.savez_compressed(os.path.join(HERE, "reference_goldens.npz"), **out)
    for k, v in out.items():
        print(f"{k:24s} {v.shape}")
