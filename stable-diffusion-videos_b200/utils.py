"""Host-side helpers mirroring stable_diffusion_videos/utils.py for the hot path's neighbours.

`slerp` (utils.py:42-66) routes CUDA tensors through the native batched kernel.  `get_timesteps_arr`
(utils.py:12-39) runs on the numpy / scipy restatement in audio.py.  `make_video_pyav` (utils.py:69-128) keeps the reference semantics (audio excerpt muxed as AAC) on top of
torchvision/PyAV when those exist and fails loudly otherwise — it never writes a mute file for an audio walk.
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
    """Audio-reactive schedule T in [0, 1] (reference utils.py:12-39).  The signal processing is this package's own
    numpy / scipy restatement (`audio.py`); librosa, when installed, is used only as the decoder
    for non-WAV inputs."""
    from . import audio

    try:
        import librosa

        y, sr = librosa.load(audio_filepath, offset=offset, duration=duration)
    except ImportError:
        y, sr = audio.load(audio_filepath, offset=offset, duration=duration)
    return audio.timesteps_from_signal(y, sr, duration, fps=fps, margin=margin, smooth=smooth)


def _video_writer():
    """torchvision.io.write_video (PyAV / ffmpeg underneath) when the installed torchvision still ships it."""
    try:
        from torchvision.io import write_video
    except ImportError as exc:
        raise RuntimeError("make_video_pyav needs torchvision.io.write_video + PyAV/ffmpeg (reference "
                           "utils.py:69-128); no H.264 / AAC encoder exists in this image and codec I/O is outside the "
                           "native hot path — call walk(make_video=False) and mux the frame%06d.png files (and the "
                           "audio excerpt) with ffmpeg") from exc
    return write_video


def make_video_pyav(frames_or_frame_dir="./frames", audio_filepath=None, fps=30, audio_offset=0, audio_duration=2,
                    sr=22050, output_filepath="output.mp4", glob_pattern="*.png"):
    """Reference utils.py:69-128: frames (a directory of images, or a (T, C, H, W) uint8 tensor) -> H.264 mp4
    (crf 10, yuv420p); with `audio_filepath`, the excerpt [audio_offset, audio_offset + audio_duration) is resampled to
    `sr`, mixed to mono and muxed as AAC.  Never drops the audio silently: without an encoder this raises."""
    write_video = _video_writer()
    from pathlib import Path

    import numpy as np
    from PIL import Image

    output_filepath = str(output_filepath)
    if isinstance(frames_or_frame_dir, (str, Path)):
        frames = [torch.from_numpy(np.asarray(Image.open(img).convert("RGB")).copy())
                  for img in sorted(Path(frames_or_frame_dir).glob(glob_pattern))]
        frames = torch.stack(frames)  # THWC already
    else:
        frames = frames_or_frame_dir.permute(0, 2, 3, 1)  # TCHW -> THWC (utils.py:102)
    opts = {"crf": "10", "pix_fmt": "yuv420p"}
    if audio_filepath:
        try:
            import librosa

            y, sr = librosa.load(audio_filepath, sr=sr, mono=True, offset=audio_offset, duration=audio_duration)
        except ImportError:
            from . import audio

            y, sr = audio.load(audio_filepath, offset=audio_offset, duration=audio_duration, sr=sr)
        write_video(output_filepath, frames, fps=fps, audio_array=torch.tensor(y).unsqueeze(0), audio_fps=sr,
                    audio_codec="aac", options=opts)
    else:
        write_video(output_filepath, frames, fps=fps, options=opts)
    return output_filepath
