"""Audio-reactive interpolation schedule without librosa.

`get_timesteps_arr` (reference utils.py:12-39) turns a music excerpt into a monotone schedule T in [0, 1] that moves
faster where the percussive energy is: load -> STFT -> harmonic/percussive separation -> iSTFT of the percussive part
-> mel power spectrogram -> per-frame max -> min-max normalise -> cumulative sum -> resample to duration * fps frames.
The reference delegates every signal-processing step to librosa, which is not installed here and is not vendored in the
reference; this module restates the published algorithms those calls implement, with librosa's documented defaults
(0.10: sr 22 050, n_fft 2048, hop 512, periodic Hann, centred frames padded with zeros, HPSS median kernels of 31 with
power-2 soft masks, 128 Slaney mel bands with Slaney area normalisation), on numpy / scipy only.

PARITY UNPINNED: there is no librosa in this image and the reference ships no golden schedule, so the restatement is
checked by construction properties only (tests/test_audio_cpu.py): STFT/iSTFT round trip, filterbank identities,
monotonicity / range / length of T, percussive emphasis on a synthetic click track.  It is a host pre-step, once per
clip; T is an INPUT of the GPU hot path.
"""
import numpy as np

SR = 22050
N_FFT = 2048
HOP = 512


# ---------------------------------------------------------------------------------------------
# loading (librosa.load(path, offset=, duration=): mono float32 at 22 050 Hz)
# ---------------------------------------------------------------------------------------------
def load(path, offset=0.0, duration=None, sr=SR):
    """WAV only (scipy.io.wavfile): PCM8/16/24/32 or float.  Other rates are resampled with a polyphase filter (librosa
    uses soxr_hq: same band limit, different filter — results agree to the filters' stop-band level, not bitwise)."""
    from scipy.io import wavfile

    native_sr, data = wavfile.read(str(path))
    start = int(round(float(offset) * native_sr))
    stop = data.shape[0] if duration is None else min(data.shape[0], start + int(round(float(duration) * native_sr)))
    data = data[start:stop]
    if data.dtype == np.uint8:
        y = (data.astype(np.float32) - 128.0) / 128.0
    elif np.issubdtype(data.dtype, np.integer):
        y = data.astype(np.float32) / float(2 ** (8 * data.dtype.itemsize - 1))
    else:
        y = data.astype(np.float32)
    if y.ndim == 2:
        y = y.mean(axis=1)
    if native_sr != sr:
        from math import gcd

        from scipy.signal import resample_poly

        g = gcd(int(sr), int(native_sr))
        y = resample_poly(y, int(sr) // g, int(native_sr) // g).astype(np.float32)
    return np.ascontiguousarray(y, dtype=np.float32), sr


# ---------------------------------------------------------------------------------------------
# STFT / iSTFT (centred, zero padded, periodic Hann, hop = n_fft / 4)
# ---------------------------------------------------------------------------------------------
def _hann(n):
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)).astype(np.float32)  # periodic ("fftbins") Hann


def stft(y, n_fft=N_FFT, hop=HOP):
    y = np.asarray(y, dtype=np.float32)
    pad = n_fft // 2
    yp = np.concatenate([np.zeros(pad, np.float32), y, np.zeros(pad, np.float32)])
    n_frames = 1 + (yp.shape[0] - n_fft) // hop
    idx = np.arange(n_fft)[:, None] + hop * np.arange(n_frames)[None, :]
    frames = yp[idx] * _hann(n_fft)[:, None]
    return np.fft.rfft(frames, axis=0).astype(np.complex64)  # [1 + n_fft/2, n_frames]


def istft(D, length=None, hop=HOP):
    n_fft = 2 * (D.shape[0] - 1)
    w = _hann(n_fft)
    frames = np.fft.irfft(D, n=n_fft, axis=0).astype(np.float32) * w[:, None]
    n_frames = D.shape[1]
    total = n_fft + hop * (n_frames - 1)
    y = np.zeros(total, np.float32)
    wss = np.zeros(total, np.float32)
    w2 = w * w
    for t in range(n_frames):  # overlap-add and the window's sum of squares, frame by frame
        y[t * hop:t * hop + n_fft] += frames[:, t]
        wss[t * hop:t * hop + n_fft] += w2
    nz = wss > np.finfo(np.float32).tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    if length is not None:
        y = y[:length] if y.shape[0] >= length else np.concatenate([y, np.zeros(length - y.shape[0], np.float32)])
    else:
        y = y[:total - n_fft]
    return y


# ---------------------------------------------------------------------------------------------
# harmonic / percussive separation (Fitzgerald 2010 median filtering, Driedger 2014 margins)
# ---------------------------------------------------------------------------------------------
def _softmask(x, x_ref, power, split_zeros):
    z = np.maximum(x, x_ref)
    bad = z < np.finfo(x.dtype).tiny
    z = np.where(bad, 1.0, z)
    mask = (x / z) ** power
    ref = (x_ref / z) ** power
    good = ~bad
    mask[good] = mask[good] / (mask[good] + ref[good])
    mask[bad] = 0.5 if split_zeros else 0.0
    return mask


def hpss(D, kernel_size=31, power=2.0, margin=1.0):
    """complex STFT -> (harmonic, percussive) complex STFTs.  Harmonic = median filter along time, percussive = along
    frequency, both on the magnitude; soft masks with exponent `power`; margin m > 1 leaves a residual in neither."""
    from scipy.ndimage import median_filter

    mag = np.abs(D).astype(np.float32)
    phase = np.where(mag > 0, D / np.maximum(mag, np.finfo(np.float32).tiny), 1.0).astype(np.complex64)
    if np.isscalar(margin):
        margin_h = margin_p = float(margin)
    else:
        margin_h, margin_p = float(margin[0]), float(margin[1])
    if margin_h < 1 or margin_p < 1:
        raise ValueError("margins must be >= 1.0")
    harm = median_filter(mag, size=(1, kernel_size), mode="reflect")
    perc = median_filter(mag, size=(kernel_size, 1), mode="reflect")
    split = margin_h == 1 and margin_p == 1
    mask_h = _softmask(harm, perc * margin_h, power, split)
    mask_p = _softmask(perc, harm * margin_p, power, split)
    return (mag * mask_h) * phase, (mag * mask_p) * phase


# ---------------------------------------------------------------------------------------------
# mel power spectrogram (Slaney scale, Slaney area normalisation)
# ---------------------------------------------------------------------------------------------
_F_SP = 200.0 / 3.0
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    return np.where(f >= _MIN_LOG_HZ, _MIN_LOG_MEL + np.log(np.maximum(f, 1e-10) / _MIN_LOG_HZ) / _LOGSTEP, f / _F_SP)


def mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    return np.where(m >= _MIN_LOG_MEL, _MIN_LOG_HZ * np.exp(_LOGSTEP * (m - _MIN_LOG_MEL)), _F_SP * m)


def mel_filterbank(sr=SR, n_fft=N_FFT, n_mels=128, fmin=0.0, fmax=None):
    fmax = sr / 2.0 if fmax is None else fmax
    fft_f = np.linspace(0.0, sr / 2.0, 1 + n_fft // 2)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_f[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]  # equal area per band
    return w.astype(np.float32)


def melspectrogram(y, sr=SR, n_fft=N_FFT, hop=HOP, n_mels=128):
    S = np.abs(stft(y, n_fft, hop)).astype(np.float32) ** 2
    return mel_filterbank(sr, n_fft, n_mels) @ S


# ---------------------------------------------------------------------------------------------
# the schedule (reference utils.py:12-39)
# ---------------------------------------------------------------------------------------------
def timesteps_from_signal(y, sr, duration, fps=30, margin=1.0, smooth=0.0):
    D = stft(y)
    _, D_percussive = hpss(D, margin=margin)
    y_percussive = istft(D_percussive, length=len(y))
    spec_raw = melspectrogram(y_percussive, sr=sr)
    spec_max = np.amax(spec_raw, axis=0)
    spec_norm = (spec_max - np.min(spec_max)) / np.ptp(spec_max)
    x_norm = np.linspace(0, spec_norm.shape[-1], spec_norm.shape[-1])
    y_norm = np.cumsum(spec_norm)
    y_norm /= y_norm[-1]
    x_resize = np.linspace(0, y_norm.shape[-1], int(duration * fps))
    T = np.interp(x_resize, x_norm, y_norm)
    return T * (1 - smooth) + np.linspace(0.0, 1.0, T.shape[0]) * smooth


def get_timesteps_arr(audio_filepath, offset, duration, fps=30, margin=1.0, smooth=0.0):
    y, sr = load(audio_filepath, offset=offset, duration=duration)
    return timesteps_from_signal(y, sr, duration, fps=fps, margin=margin, smooth=smooth)
