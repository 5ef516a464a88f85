"""Host-side helpers mirroring stable_diffusion_videos/utils.py for the hot path's neighbours.

`slerp` (utils.py:42-66) routes CUDA tensors through the native batched kernel.  `get_timesteps_arr`
(utils.py:12-39) uses librosa when installed and the numpy / scipy restatement in audio.py otherwise (SURVEY.md §8f
row 2).  `make_video_pyav` (utils.py:69-128, torchvision/PyAV) is OUT of scope: it delegates or fails loudly.
"""
import torch

from . import _native


def slerp(t, v0, v1, DOT_THRESHOLD=0.9995):
    """spherical interpolation of two CUDA tensors (reference utils.py:42); fp32 math, rounded once."""
    if not (isinstance(v0, torch.Tensor) and v0.is_cuda):
        raise _native.SdwError("native slerp takes CUDA tensors (the reference's host numpy round-trip is what it "
                               "replaces); there is no CPU path in this package")
    tt = torch.tensor([float(t)], dtype=torch.float32, device=v0.device)
    e = torch.zeros((1, 8), dtype=v0.dtype, device=v0.device)
    out, _ = _native.slerp_lerp_batch(v0.reshape(1, -1), v1.reshape(1, -1), e, e, tt, DOT_THRESHOLD)
    return out.reshape(v0.shape)


def get_timesteps_arr(audio_filepath, offset, duration, fps=30, margin=1.0, smooth=0.0):
    """Audio-reactive schedule T in [0, 1] (reference utils.py:12-39).  Uses librosa when it is installed (bit-for-bit
    the reference's calls); otherwise the numpy / scipy restatement in `audio.py` (WAV input)."""
    try:
        import librosa
    except ImportError:
        from . import audio

        return audio.get_timesteps_arr(audio_filepath, offset, duration, fps=fps, margin=margin, smooth=smooth)
    import numpy as np

    y, sr = librosa.load(audio_filepath, offset=offset, duration=duration)
    D = librosa.stft(y, n_fft=2048, hop_length=512)
    D_harmonic, D_percussive = librosa.decompose.hpss(D, margin=margin)
    y_percussive = librosa.istft(D_percussive, length=len(y))
    spec_raw = librosa.feature.melspectrogram(y=y_percussive, sr=sr)
    spec_max = np.amax(spec_raw, axis=0)
    spec_norm = (spec_max - np.min(spec_max)) / np.ptp(spec_max)
    x_norm = np.linspace(0, spec_norm.shape[-1], spec_norm.shape[-1])
    y_norm = np.cumsum(spec_norm)
    y_norm /= y_norm[-1]
    x_resize = np.linspace(0, y_norm.shape[-1], int(duration * fps))
    T = np.interp(x_resize, x_norm, y_norm)
    return T * (1 - smooth) + np.linspace(0.0, 1.0, T.shape[0]) * smooth


def make_video_pyav(frames_or_frame_dir="./frames", audio_filepath=None, fps=30, audio_offset=0, audio_duration=2,
                    sr=22050, output_filepath="output.mp4", glob_pattern="*.png"):
    try:
        from torchvision.io import write_video
    except ImportError as exc:
        raise RuntimeError("make_video_pyav needs torchvision.io.write_video + PyAV/ffmpeg (reference "
                           "utils.py:69-128); codec I/O is outside the native hot path — call walk(make_video=False) "
                           "and mux the frame%06d.png files with ffmpeg") from exc
    from pathlib import Path

    import numpy as np
    from PIL import Image

    output_filepath = str(output_filepath)
    if isinstance(frames_or_frame_dir, (str, Path)):
        frames = None
        for img in sorted(Path(frames_or_frame_dir).glob(glob_pattern)):
            frame = torch.from_numpy(np.asarray(Image.open(img).convert("RGB"))).unsqueeze(0)
            frames = frame if frames is None else torch.cat([frames, frame])
    else:
        frames = frames_or_frame_dir
    write_video(output_filepath, frames, fps=fps, options={"crf": "10", "pix_fmt": "yuv420p"})
    return output_filepath
