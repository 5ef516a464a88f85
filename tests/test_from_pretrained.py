"""StableDiffusionWalkPipeline.from_pretrained (reference P:841-858) on a fabricated LOCAL diffusers-layout checkpoint
(tests/_fake_checkpoint.py: tiny UNet / VAE / CLIP text tower, a 74-token CLIP tokenizer, scheduler_config.json).
CPU: the configuration checks that run before any weight reaches the GPU.  GPU: the loaded pipeline must produce the same
frames as one assembled by hand from the same state dicts — config parsing, safetensors reading, the old VAE attention
key names (query / key / value / proj_attn) and the CLIP tower hand-over all sit on that path."""
import json
import os

import numpy as np
import pytest
import torch

import _fake_checkpoint as fc


def _pipeline_cls():
    from stable_diffusion_videos_b200.pipeline import StableDiffusionWalkPipeline
    return StableDiffusionWalkPipeline


def test_from_pretrained_needs_a_local_directory(tmp_path):
    with pytest.raises(FileNotFoundError):
        _pipeline_cls().from_pretrained(str(tmp_path / "nope"))


@pytest.mark.parametrize("sched,extra,exc", [
    ("EulerAncestralDiscreteScheduler", {}, NotImplementedError),      # stochastic sampler: no native plan
    ("PNDMScheduler", {"trained_betas": [0.1, 0.2]}, NotImplementedError),
    ("DDIMScheduler", {"set_alpha_to_one": True}, NotImplementedError),
    ("PNDMScheduler", {"skip_prk_steps": False}, NotImplementedError),
    ("DDIMScheduler", {"prediction_type": "sample"}, NotImplementedError),
    ("PNDMScheduler", {"beta_schedule": "linear"}, ValueError),
])
def test_scheduler_config_is_validated_not_ignored(tmp_path, sched, extra, exc):
    fc.write_checkpoint(str(tmp_path), scheduler=sched, scheduler_extra=extra, with_weights=False)
    with pytest.raises(exc):
        _pipeline_cls().from_pretrained(str(tmp_path))


def test_fake_checkpoint_layout(tmp_path):
    fc.write_checkpoint(str(tmp_path), with_weights=False)
    assert json.loads((tmp_path / "unet" / "config.json").read_text())["cross_attention_dim"] == 64
    assert sorted(os.listdir(tmp_path)) == ["scheduler", "tokenizer", "unet", "vae"]


@pytest.mark.gpu
def test_from_pretrained_equals_hand_assembled_pipeline(tmp_path):
    from transformers import CLIPTokenizer

    from _helpers import TINY_UNET, TINY_VAE, product_cfgs
    from stable_diffusion_videos_b200.clip import NativeCLIPTextEncoder
    from stable_diffusion_videos_b200.pipeline import NativeUNet, NativeVAE
    from stable_diffusion_videos_b200.schedulers import PNDMScheduler

    unet, vae, te = fc.write_checkpoint(str(tmp_path))
    cls = _pipeline_cls()
    pipe = cls.from_pretrained(str(tmp_path), torch_dtype=torch.float16, safety_checker=None).to("cuda")
    assert pipe.tiled is False and pipe.tokenizer.model_max_length == 77
    assert type(pipe.scheduler).__name__ == "PNDMScheduler" and pipe.unet.config.sample_size == TINY_UNET.sample_size

    ucfg, vcfg = product_cfgs(TINY_UNET, TINY_VAE)
    ref = cls(NativeVAE(vcfg, {k: v.half() for k, v in vae.state_dict().items()}),
              NativeCLIPTextEncoder.from_hf_model(te), CLIPTokenizer.from_pretrained(str(tmp_path / "tokenizer")),
              NativeUNet(ucfg, {k: v.half() for k, v in unet.state_dict().items()}), PNDMScheduler()).to("cuda")

    # text tower: native vs the transformers module the checkpoint was written from (fp32, CPU)
    ids = pipe.tokenizer(["ab cd"], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
    with torch.no_grad():
        want = te(ids)[0]
    got = pipe.embed_text("ab cd").float().cpu()
    assert got.shape == want.shape and (got - want).abs().max() <= 2e-2 * want.abs().max()  # tests/test_clip_gpu.py's bound

    g = torch.Generator().manual_seed(7)
    lat = torch.randn(2, 4, 8, 8, generator=g).half().cuda()
    kw = dict(prompt=["ab cd", "xyz 12"], height=64, width=64, num_inference_steps=3, guidance_scale=7.5, latents=lat,
              output_type="numpy")
    a, b = pipe(**kw)["images"], ref(**kw)["images"]
    assert a.shape == b.shape and a.shape[0] == 2 and np.isfinite(a).all()
    assert np.array_equal(a, b)  # same weights through both constructors: bit-identical frames
