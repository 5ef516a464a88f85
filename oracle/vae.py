"""ORACLE (test infrastructure only).  CPU fp32 restatement of `diffusers.AutoencoderKL.decode`, called by the
reference at stable_diffusion_pipeline.py:433 (`self.vae.decode(latents).sample`).  Architecture restated from
the published SD-1.x `vae/config.json`; names follow the diffusers state-dict.  PARITY UNPINNED
(see oracle/__init__.py); pinned by the published parameter count 49,490,199 (decoder + post_quant_conv).
"""
from dataclasses import dataclass
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .unet import ResnetBlock2D, Upsample2D


@dataclass
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    norm_num_groups: int = 32

    @staticmethod
    def tiny(ch=(32, 64), groups=8):
        return VAEConfig(block_out_channels=ch, norm_num_groups=groups, layers_per_block=1)


class AttentionBlock(nn.Module):
    """single-head spatial self-attention (d = C), GN eps 1e-6, residual."""

    def __init__(self, ch, groups):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.to_q = nn.Linear(ch, ch)
        self.to_k = nn.Linear(ch, ch)
        self.to_v = nn.Linear(ch, ch)
        self.to_out = nn.ModuleList([nn.Linear(ch, ch)])

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x).permute(0, 2, 3, 1).reshape(b, h * w, c)
        q, k, v = self.to_q(y), self.to_k(y), self.to_v(y)
        s = torch.softmax(q @ k.transpose(-1, -2) * c ** -0.5, dim=-1)
        o = self.to_out[0](s @ v)
        return x + o.reshape(b, h, w, c).permute(0, 3, 1, 2)


class VAEMidBlock(nn.Module):
    def __init__(self, ch, groups):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, 0, groups, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([AttentionBlock(ch, groups)])

    def forward(self, h):
        return self.resnets[1](self.attentions[0](self.resnets[0](h)))


class UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, n, groups, add_up):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if j == 0 else cout, cout, 0, groups, 1e-6)
                                      for j in range(n)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class Decoder(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        ch = cfg.block_out_channels
        self.conv_in = nn.Conv2d(cfg.latent_channels, ch[-1], 3, padding=1)
        self.mid_block = VAEMidBlock(ch[-1], cfg.norm_num_groups)
        rev = list(reversed(ch))
        self.up_blocks = nn.ModuleList()
        cout = rev[0]
        for i in range(len(ch)):
            cin, cout = cout, rev[i]
            self.up_blocks.append(UpDecoderBlock(cin, cout, cfg.layers_per_block + 1, cfg.norm_num_groups,
                                                 add_up=i < len(ch) - 1))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-6)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z))
        for blk in self.up_blocks:
            h = blk(h)
        return self.conv_out(F.silu(self.conv_norm_out(h)))


class AutoencoderKLDecoder(nn.Module):
    """`post_quant_conv` + `decoder` of AutoencoderKL (the only half the hot path uses)."""

    def __init__(self, cfg: VAEConfig = None):
        super().__init__()
        cfg = cfg or VAEConfig()
        self.cfg = cfg
        self.post_quant_conv = nn.Conv2d(cfg.latent_channels, cfg.latent_channels, 1)
        self.decoder = Decoder(cfg)

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))
