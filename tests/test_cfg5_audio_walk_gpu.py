"""BASELINE.json configs[4] as one piece (tests/test_pipeline.py:53-68, examples/make_music_video.py:43-55): an
audio-driven walk — K-LMS, guidance 15, smooth 0.2, margin 1.0, batch 12 — through walk() on the GPU.

Two legs: (1) walk() with an audio file end to end (a synthetic click track written here: the reference's choice.wav does
not travel to the GPU box), checking that the schedule it derives is the one the frames follow; (2) frames rendered along
the COMMITTED schedule of choice.wav (tests/golden/cfg5_choice_T.npy) through make_clip_frames, against the oracle fed the
same T.  Per-frame numerics of LMS-50 / guidance 15 at full size are pinned by tests/test_golden_full_gpu.py."""
import os

import numpy as np
import pytest
import torch
from PIL import Image

from _helpers import TINY_UNET, TINY_VAE, make_oracle, product_cfgs

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lms_pipe():
    from stable_diffusion_videos_b200.pipeline import (NativeUNet, NativeVAE, StableDiffusionWalkPipeline,
                                                       SyntheticTextEncoder, SyntheticTokenizer)
    from stable_diffusion_videos_b200.schedulers import LMSDiscreteScheduler

    unet, vae = make_oracle(TINY_UNET, TINY_VAE)
    ucfg, vcfg = product_cfgs(TINY_UNET, TINY_VAE)
    usd = {k: v.half() for k, v in unet.state_dict().items()}
    vsd = {k: v.half() for k, v in vae.state_dict().items()}
    pipe = StableDiffusionWalkPipeline(NativeVAE(vcfg, vsd), SyntheticTextEncoder(TINY_UNET.cross_attention_dim),
                                       SyntheticTokenizer(), NativeUNet(ucfg, usd), LMSDiscreteScheduler()).to("cuda")
    return pipe, unet, vae


def _click_track(path, sr=22050, seconds=1.0):
    from scipy.io import wavfile

    rng = np.random.default_rng(0)
    t = np.arange(int(sr * seconds)) / sr
    y = 0.2 * np.sin(2 * np.pi * 220.0 * t)
    for c in (0.2, 0.6, 0.7, 0.8):
        i = int(c * sr)
        y[i:i + 256] += rng.standard_normal(256) * np.hanning(256) * 0.9
    wavfile.write(str(path), sr, (y * 32767).astype(np.int16))


def test_audio_walk_end_to_end(lms_pipe, tmp_path):
    from stable_diffusion_videos_b200.utils import get_timesteps_arr

    pipe, _, _ = lms_pipe
    wav = tmp_path / "clicks.wav"
    _click_track(wav)
    pipe.walk(["0", "1"], seeds=[42, 1337], num_interpolation_steps=[12], fps=12, audio_filepath=str(wav),
              audio_start_sec=0, batch_size=12, num_inference_steps=6, guidance_scale=15, margin=1.0, smooth=0.2,
              output_dir=str(tmp_path), name="mv", height=64, width=64, make_video=False)
    files = sorted((tmp_path / "mv" / "mv_000000").glob("*.png"))
    assert [f.name for f in files] == [f"frame{i:06d}.png" for i in range(12)]
    T = get_timesteps_arr(str(wav), offset=0, duration=1.0, fps=12, margin=1.0, smooth=0.2)
    assert T.shape == (12,) and np.all(np.diff(T) >= 0)
    # the frames follow T: rendering the same clip with that explicit T gives the same files
    pipe.make_clip_frames("0", "1", 42, 1337, num_interpolation_steps=12, save_path=tmp_path / "explicit", T=T,
                          batch_size=12, num_inference_steps=6, guidance_scale=15, height=64, width=64)
    for f in files:
        a, b = np.asarray(Image.open(f)), np.asarray(Image.open(tmp_path / "explicit" / f.name))
        assert np.array_equal(a, b), f.name


def test_frames_along_the_choice_wav_schedule_match_oracle(lms_pipe, tmp_path):
    from oracle.pipeline import generate_inputs, sample_frames, to_uint8
    from oracle.schedulers import make_scheduler

    pipe, unet, vae = lms_pipe
    T = np.load(os.path.join(HERE, "golden", "cfg5_choice_T.npy"))[::25]  # 12 of the 300 frames of the 10 s clip
    pipe.make_clip_frames("0", "1", 42, 1337, num_interpolation_steps=12, save_path=tmp_path / "c", T=T, batch_size=12,
                          num_inference_steps=6, guidance_scale=15, height=64, width=64)
    got = np.stack([np.asarray(Image.open(f)) for f in sorted((tmp_path / "c").glob("*.png"))])
    ea, eb = pipe.embed_text("0").float().cpu(), pipe.embed_text("1").float().cpu()
    la = pipe.init_noise(42, (1, 4, 8, 8), torch.float16).float().cpu()
    lb = pipe.init_noise(1337, (1, 4, 8, 8), torch.float16).float().cpu()
    unc = pipe._uncond([""]).float().cpu()
    (_, e, z), = list(generate_inputs(ea, eb, la, lb, T, 12))
    ref = to_uint8(sample_frames(unet, vae, make_scheduler("lms"), z, e, unc, 6, 15.0))
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.mean() <= 1.0 and (d <= 2).mean() >= 0.99 and d.max() <= 8, (d.mean(), d.max())
