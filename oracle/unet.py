"""ORACLE (test infrastructure only — never imported by the product path).

CPU fp32 restatement of `diffusers.UNet2DConditionModel.forward`, the model the reference calls at
stable_diffusion_videos/stable_diffusion_pipeline.py:418 (`self.unet(latent_model_input, t,
encoder_hidden_states=text_embeddings).sample`).  diffusers is an un-vendored, un-pinned dependency
of the reference (pyproject.toml:14) and is NOT installable here, so this file restates its published
architecture (SD-1.x / SD-2.x `unet/config.json`); module / parameter names follow the diffusers
state-dict so real checkpoints load unchanged.  PARITY UNPINNED: the reference holds
no golden vectors for this path (tests/test_pipeline.py:50,68,81 assert only file existence); the pins
this repo adds are the exact published parameter counts (859,520,964 / 865,910,724) and key sets.
"""
import math
from dataclasses import dataclass
from typing import Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


@dataclass
class UNetConfig:
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: Union[int, Tuple[int, ...]] = 8  # diffusers' (mis)name for the NUMBER of heads
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    sample_size: int = 64
    use_linear_projection: bool = False
    flip_sin_to_cos: bool = True
    freq_shift: int = 0

    @staticmethod
    def sd14():
        return UNetConfig()

    @staticmethod
    def sd21():
        return UNetConfig(attention_head_dim=(5, 10, 20, 20), cross_attention_dim=1024,
                          use_linear_projection=True, sample_size=96)

    @staticmethod
    def tiny(ch=(32, 64, 64, 64), heads=4, cross=64, groups=8, sample=8):
        return UNetConfig(block_out_channels=ch, attention_head_dim=heads, cross_attention_dim=cross,
                          norm_num_groups=groups, sample_size=sample)

    def heads(self, level):
        a = self.attention_head_dim
        return a if isinstance(a, int) else a[level]


def timestep_embedding(t, dim, flip_sin_to_cos=True, freq_shift=0, max_period=10000):
    """diffusers `Timesteps` / `get_timestep_embedding`: cat[cos(t f), sin(t f)]."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32) / (half - freq_shift)
    emb = t.float()[:, None] * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


class ResnetBlock2D(nn.Module):
    def __init__(self, cin, cout, temb_ch, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        if temb_ch:
            self.time_emb_proj = nn.Linear(temb_ch, cout)
        else:
            self.time_emb_proj = None
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim, heads):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(query_dim, query_dim, bias=False)
        self.to_k = nn.Linear(context_dim, query_dim, bias=False)
        self.to_v = nn.Linear(context_dim, query_dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(query_dim, query_dim)])

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, c = x.shape
        h = self.heads
        q = self.to_q(x).reshape(b, n, h, c // h).transpose(1, 2)
        k = self.to_k(context).reshape(b, -1, h, c // h).transpose(1, 2)
        v = self.to_v(context).reshape(b, -1, h, c // h).transpose(1, 2)
        s = torch.softmax(q @ k.transpose(-1, -2) * (c // h) ** -0.5, dim=-1)
        o = (s @ v).transpose(1, 2).reshape(b, n, c)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Linear(cin, cout * 2)

    def forward(self, x):
        a, g = self.proj(x).chunk(2, dim=-1)
        return a * F.gelu(g)


class FeedForward(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, context_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = CrossAttention(dim, dim, heads)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = CrossAttention(dim, context_dim, heads)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, x, context):
        x = x + self.attn1(self.norm1(x))
        x = x + self.attn2(self.norm2(x), context)
        x = x + self.ff(self.norm3(x))
        return x


class Transformer2DModel(nn.Module):
    def __init__(self, ch, heads, context_dim, groups, linear_proj):
        super().__init__()
        self.linear_proj = linear_proj
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Linear(ch, ch) if linear_proj else nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(ch, heads, context_dim)])
        self.proj_out = nn.Linear(ch, ch) if linear_proj else nn.Conv2d(ch, ch, 1)

    def forward(self, x, context):
        b, c, hh, ww = x.shape
        res = x
        h = self.norm(x)
        if self.linear_proj:
            h = self.proj_in(h.permute(0, 2, 3, 1).reshape(b, hh * ww, c))
        else:
            h = self.proj_in(h).permute(0, 2, 3, 1).reshape(b, hh * ww, c)
        for blk in self.transformer_blocks:
            h = blk(h, context)
        if self.linear_proj:
            h = self.proj_out(h).reshape(b, hh, ww, c).permute(0, 3, 1, 2)
        else:
            h = self.proj_out(h.reshape(b, hh, ww, c).permute(0, 3, 1, 2))
        return h + res


class Downsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class DownBlock(nn.Module):
    def __init__(self, cfg, level, cin, cout, temb_ch, has_attn, add_down):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if j == 0 else cout, cout, temb_ch, cfg.norm_num_groups, cfg.norm_eps)
            for j in range(cfg.layers_per_block)])
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, cfg.heads(level), cfg.cross_attention_dim, cfg.norm_num_groups,
                                   cfg.use_linear_projection) for _ in range(cfg.layers_per_block)])
        else:
            self.attentions = None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)]) if add_down else None

    def forward(self, h, temb, ctx):
        outs = []
        for j, res in enumerate(self.resnets):
            h = res(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, cfg, ch, temb_ch):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(ch, ch, temb_ch, cfg.norm_num_groups, cfg.norm_eps)
                                      for _ in range(2)])
        self.attentions = nn.ModuleList([Transformer2DModel(ch, cfg.heads(len(cfg.block_out_channels) - 1),
                                                            cfg.cross_attention_dim, cfg.norm_num_groups,
                                                            cfg.use_linear_projection)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    def __init__(self, cfg, level, cin, cout, prev, temb_ch, has_attn, add_up):
        super().__init__()
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        for j in range(n):
            skip = cin if j == n - 1 else cout
            rin = prev if j == 0 else cout
            self.resnets.append(ResnetBlock2D(rin + skip, cout, temb_ch, cfg.norm_num_groups, cfg.norm_eps))
        if has_attn:
            self.attentions = nn.ModuleList([
                Transformer2DModel(cout, cfg.heads(level), cfg.cross_attention_dim, cfg.norm_num_groups,
                                   cfg.use_linear_projection) for _ in range(n)])
        else:
            self.attentions = None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)]) if add_up else None

    def forward(self, h, skips, temb, ctx):
        for j, res in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = res(h, temb)
            if self.attentions is not None:
                h = self.attentions[j](h, ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig = None):
        super().__init__()
        cfg = cfg or UNetConfig()
        self.cfg = cfg
        ch = cfg.block_out_channels
        temb_ch = ch[0] * 4
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, padding=1)
        self.time_embedding = TimestepEmbedding(ch[0], temb_ch)
        nlev = len(ch)
        self.down_blocks = nn.ModuleList()
        cout = ch[0]
        for i in range(nlev):
            cin, cout = cout, ch[i]
            last = i == nlev - 1
            self.down_blocks.append(DownBlock(cfg, i, cin, cout, temb_ch, has_attn=not last, add_down=not last))
        self.mid_block = MidBlock(cfg, ch[-1], temb_ch)
        self.up_blocks = nn.ModuleList()
        rev = list(reversed(ch))
        cout = rev[0]
        for i in range(nlev):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, nlev - 1)]
            last = i == nlev - 1
            self.up_blocks.append(UpBlock(cfg, nlev - 1 - i, cin, cout, prev, temb_ch, has_attn=i > 0,
                                          add_up=not last))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=cfg.norm_eps)
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, padding=1)

    def time_embed(self, t, batch):
        t = torch.as_tensor(t, dtype=torch.float32).reshape(-1)
        if t.numel() == 1:
            t = t.expand(batch)
        emb = timestep_embedding(t, self.cfg.block_out_channels[0], self.cfg.flip_sin_to_cos, self.cfg.freq_shift)
        return self.time_embedding(emb.to(self.conv_in.weight.dtype))

    def forward(self, x, t, encoder_hidden_states):
        temb = self.time_embed(t, x.shape[0])
        h = self.conv_in(x)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, temb, encoder_hidden_states)
            skips.extend(outs)
        h = self.mid_block(h, temb, encoder_hidden_states)
        for blk in self.up_blocks:
            h = blk(h, skips, temb, encoder_hidden_states)
        return self.conv_out(F.silu(self.conv_norm_out(h)))
