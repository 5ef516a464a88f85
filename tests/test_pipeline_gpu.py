"""(tiny VAE: 2 levels, so a 64-px request — latent 8x8, the reference hard-codes //8 at P:368 — decodes to 16x16)
The drop-in surface on the GPU: walk() / make_clip_frames() / generate_inputs() / __call__ of the
native StableDiffusionWalkPipeline — file layout, resume, error behaviour, and frame parity with the oracle's
restatement of the same control flow (reference scenarios: tests/test_pipeline.py:41-50)."""
import json

import numpy as np
import pytest
import torch
from PIL import Image

from _helpers import TINY_UNET, TINY_VAE, make_oracle, product_cfgs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pipe_and_oracle():
    from stable_diffusion_videos_b200.pipeline import (NativeUNet, NativeVAE, StableDiffusionWalkPipeline,
                                                       SyntheticTextEncoder, SyntheticTokenizer)
    from stable_diffusion_videos_b200.schedulers import PNDMScheduler

    unet, vae = make_oracle(TINY_UNET, TINY_VAE)
    ucfg, vcfg = product_cfgs(TINY_UNET, TINY_VAE)
    usd = {k: v.half() for k, v in unet.state_dict().items()}
    vsd = {k: v.half() for k, v in vae.state_dict().items()}
    pipe = StableDiffusionWalkPipeline(NativeVAE(vcfg, vsd), SyntheticTextEncoder(TINY_UNET.cross_attention_dim),
                                       SyntheticTokenizer(), NativeUNet(ucfg, usd), PNDMScheduler()).to("cuda")
    return pipe, unet, vae


def test_walk_writes_reference_layout_and_matches_oracle(pipe_and_oracle, tmp_path):
    from oracle.pipeline import walk_frames
    from oracle.schedulers import make_scheduler

    pipe, unet, vae = pipe_and_oracle
    out = pipe.walk(["0", "1", "2"], seeds=[42, 1337, 2022], num_interpolation_steps=[3, 3], output_dir=str(tmp_path),
                    name="run", fps=3, num_inference_steps=4, height=64, width=64, batch_size=2, make_video=False)
    assert out is None  # make_video=False (stable_diffusion_pipeline.py:578, 797)
    root = tmp_path / "run"
    cfg = json.loads((root / "prompt_config.json").read_text())
    assert cfg["prompts"] == ["0", "1", "2"] and cfg["num_interpolation_steps"] == [3, 3] and cfg["height"] == 64
    files = sorted(p.relative_to(root).as_posix() for p in root.glob("**/*.png"))
    assert files == [f"run_{i:06d}/frame{j:06d}.png" for i in range(2) for j in range(3)]
    got = np.stack([np.asarray(Image.open(root / f)) for f in files])
    # the oracle's CPU generator differs from the CUDA generator init_noise uses (P:832-837): feed it the same noise
    unc = pipe._uncond([""]).float().cpu()
    frames = []
    for k in range(2):
        ea, eb = pipe.embed_text(str(k)).float().cpu(), pipe.embed_text(str(k + 1)).float().cpu()
        la = pipe.init_noise([42, 1337, 2022][k], (1, 4, 8, 8), torch.float16).float().cpu()
        lb = pipe.init_noise([42, 1337, 2022][k + 1], (1, 4, 8, 8), torch.float16).float().cpu()
        from oracle.pipeline import generate_inputs, sample_frames, to_uint8

        for _, e, z in generate_inputs(ea, eb, la, lb, np.linspace(0.0, 1.0, 3), 3):
            frames.append(to_uint8(sample_frames(unet, vae, make_scheduler("pndm"), z, e, unc, 4, 7.5)))
    ref = np.concatenate(frames)
    d = np.abs(got.astype(np.int32) - ref.astype(np.int32))
    assert d.mean() <= 1.0 and (d <= 2).mean() >= 0.99 and d.max() <= 8, (d.mean(), d.max())


def test_walk_resume_skips_finished_frames(pipe_and_oracle, tmp_path):
    pipe, _, _ = pipe_and_oracle
    kw = dict(output_dir=str(tmp_path), name="r", num_inference_steps=2, height=64, width=64, make_video=False)
    pipe.walk(["0", "1"], seeds=[1, 2], num_interpolation_steps=4, batch_size=2, **kw)
    clip = tmp_path / "r" / "r_000000"
    (clip / "frame000002.png").unlink()
    (clip / "frame000003.png").unlink()
    first = (clip / "frame000000.png").stat().st_mtime_ns
    pipe.walk(resume=True, batch_size=2, **kw)  # everything else is re-read from prompt_config.json (P:715-729)
    assert sorted(p.name for p in clip.glob("*.png")) == [f"frame{i:06d}.png" for i in range(4)]
    assert (clip / "frame000000.png").stat().st_mtime_ns == first


def test_call_surface_and_errors(pipe_and_oracle):
    from stable_diffusion_videos_b200 import _native

    pipe, _, _ = pipe_and_oracle
    out = pipe(prompt=["0", "1"], height=64, width=64, num_inference_steps=2, output_type="numpy")
    assert out["images"].shape == (2, 16, 16, 3) and out.images.dtype == np.float32
    assert 0.0 <= out.images.min() and out.images.max() <= 1.0 and out.nsfw_content_detected is None
    imgs, nsfw = pipe(prompt="0", height=64, width=64, num_inference_steps=2, return_dict=False)
    assert isinstance(imgs[0], Image.Image) and imgs[0].size == (16, 16) and nsfw is None
    with pytest.raises(ValueError, match="divisible by 8"):
        pipe(prompt="0", height=20, width=64)
    with pytest.raises(ValueError, match="callback_steps"):
        pipe(prompt="0", height=64, width=64, callback_steps=0)
    with pytest.raises(ValueError, match="`prompt` has to be of type"):
        pipe(prompt=3, height=64, width=64)
    with pytest.raises(ValueError, match="Unexpected latents shape"):
        pipe(prompt="0", height=64, width=64, latents=torch.zeros(1, 4, 3, 3, device="cuda"))
    with pytest.raises(ValueError, match="Unexpected T shape"):
        pipe.make_clip_frames("0", "1", 1, 2, num_interpolation_steps=3, T=np.linspace(0, 1, 4), height=64, width=64)
    with pytest.raises(_native.SdwError):
        pipe.to("cpu")  # no CPU path: the product fails loudly
    assert pipe.unet.in_channels == 4 and pipe.vae_scale_factor == 2 and pipe.tokenizer.model_max_length == 77


def test_generate_inputs_batches_like_the_reference(pipe_and_oracle):
    pipe, _, _ = pipe_and_oracle
    T = np.linspace(0.0, 1.0, 5)
    batches = list(pipe.generate_inputs("0", "1", 42, 1337, (1, 4, 8, 8), T, 2))
    assert [b[0] for b in batches] == [0, 1, 2]
    assert [b[1].shape[0] for b in batches] == [2, 2, 1] and batches[0][2].shape == (2, 4, 8, 8)
    ea, eb = pipe.embed_text("0"), pipe.embed_text("1")
    la = pipe.init_noise(42, (1, 4, 8, 8), ea.dtype)
    assert torch.equal(batches[0][2][0], la[0]) and torch.equal(batches[0][1][0], ea[0])  # t = 0 endpoints exact
    assert torch.equal(batches[2][1][0], eb[0])


def test_callback_sees_every_step_and_does_not_change_the_frames(pipe_and_oracle):
    """per-step callback(i, t, latents) (stable_diffusion_pipeline.py:429-430): the segmented eager sampler must report the
    scheduler's timesteps in order, hand out the evolving latents, and end on the same frames as the fused graph."""
    from oracle.pipeline import sample_frames
    from oracle.schedulers import make_scheduler

    pipe, unet, vae = pipe_and_oracle
    g = torch.Generator(device="cuda").manual_seed(3)
    lat = torch.randn(2, 4, 8, 8, generator=g, device="cuda", dtype=torch.float16)
    emb = torch.cat([pipe.embed_text("0"), pipe.embed_text("1")])
    seen = []
    a = pipe(text_embeddings=emb, latents=lat, height=64, width=64, num_inference_steps=4, output_type="numpy",
             callback=lambda i, t, x: seen.append((i, int(t), x.float().cpu())), callback_steps=2)
    b = pipe(text_embeddings=emb, latents=lat, height=64, width=64, num_inference_steps=4, output_type="numpy")
    assert [s[0] for s in seen] == [0, 2, 4] and [s[1] for s in seen] == [751, 501, 1]  # PNDM-4: 751 501 501 251 1
    assert np.abs(a.images - b.images).max() <= 1e-6
    # the latents the callback sees are the oracle's after the same number of steps
    ref = []
    sample_frames(unet, vae, make_scheduler("pndm"), lat.float().cpu(), emb.float().cpu(), pipe._uncond([""]).float().cpu(),
                  4, 7.5, callback=lambda i, t, x: ref.append(x.clone()))
    for (i, _, x) in seen:
        assert float((x - ref[i]).norm() / ref[i].norm()) <= 1e-2


def test_per_sample_negative_prompts(pipe_and_oracle):
    """a list `negative_prompt` (one per prompt, P:331-358): each sample is guided away from ITS negative embedding —
    sample k of the batch call equals the single-sample call with negative prompt k."""
    pipe, _, _ = pipe_and_oracle
    g = torch.Generator(device="cuda").manual_seed(5)
    lat = torch.randn(2, 4, 8, 8, generator=g, device="cuda", dtype=torch.float16)
    kw = dict(height=64, width=64, num_inference_steps=3, output_type="numpy")
    both = pipe(prompt=["0", "1"], negative_prompt=["7", "8"], latents=lat, **kw).images
    for k, neg in enumerate(["7", "8"]):
        one = pipe(prompt=str(k), negative_prompt=neg, latents=lat[k:k + 1], **kw).images
        assert np.abs(both[k] - one[0]).max() <= 2e-3
    assert np.abs(both[0] - pipe(prompt="0", negative_prompt="8", latents=lat[:1], **kw).images[0]).max() > 1e-2
    with pytest.raises(ValueError, match="batch size"):
        pipe(prompt=["0", "1"], negative_prompt=["7"], **kw)
