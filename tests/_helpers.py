"""Shared test helpers: build oracle modules with fp16-rounded weights, and matching product configs."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle.unet import UNet2DConditionModel, UNetConfig as OUNetConfig  # noqa: E402
from oracle.vae import AutoencoderKLDecoder, VAEConfig as OVAEConfig  # noqa: E402


def round_to_f16_(module):
    """the same fp16-rounded weights feed oracle (fp32 math) and native path."""
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.half().float())
    return module


def make_oracle(unet_cfg: OUNetConfig, vae_cfg: OVAEConfig, seed=0, out_gain=1.0):
    torch.manual_seed(seed)
    unet = UNet2DConditionModel(unet_cfg).eval()
    vae = AutoencoderKLDecoder(vae_cfg).eval()
    if out_gain != 1.0:
        with torch.no_grad():
            unet.conv_out.weight.mul_(out_gain)
    round_to_f16_(unet)
    round_to_f16_(vae)
    return unet, vae


def product_cfgs(unet_cfg: OUNetConfig, vae_cfg: OVAEConfig):
    from stable_diffusion_videos_b200.configs import UNetConfig, VAEConfig

    u = UNetConfig(in_channels=unet_cfg.in_channels, out_channels=unet_cfg.out_channels,
                   block_out_channels=tuple(unet_cfg.block_out_channels), layers_per_block=unet_cfg.layers_per_block,
                   attention_head_dim=unet_cfg.attention_head_dim, cross_attention_dim=unet_cfg.cross_attention_dim,
                   norm_num_groups=unet_cfg.norm_num_groups, norm_eps=unet_cfg.norm_eps,
                   sample_size=unet_cfg.sample_size, use_linear_projection=unet_cfg.use_linear_projection)
    v = VAEConfig(latent_channels=vae_cfg.latent_channels, out_channels=vae_cfg.out_channels,
                  block_out_channels=tuple(vae_cfg.block_out_channels), layers_per_block=vae_cfg.layers_per_block,
                  norm_num_groups=vae_cfg.norm_num_groups)
    return u, v


TINY_UNET = OUNetConfig(block_out_channels=(32, 64, 64, 64), attention_head_dim=4, cross_attention_dim=64,
                        norm_num_groups=8, sample_size=8)
TINY_VAE = OVAEConfig(block_out_channels=(32, 64), layers_per_block=1, norm_num_groups=8)
# SD-1.4 block structure with narrower but kernel-relevant widths (exercises BLOCK_N 64/128/160, d = 40/80)
MID_UNET = OUNetConfig(block_out_channels=(320, 640, 640, 640), attention_head_dim=8, cross_attention_dim=768,
                       norm_num_groups=32, sample_size=16)
MID_VAE = OVAEConfig(block_out_channels=(64, 128, 128), layers_per_block=1, norm_num_groups=32)


def set_tiled(*modules):
    """the reference's `tiled=True` (stable_diffusion_pipeline.py:841-858): every nn.Conv2d pads circularly"""
    for m in modules:
        for sub in m.modules():
            if isinstance(sub, torch.nn.Conv2d) and sub.padding != (0, 0):
                sub.padding_mode = "circular"
    return modules
