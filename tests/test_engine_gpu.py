"""GPU parity tests of the native engine (UNet forward, VAE decode, full sampler) against the CPU fp32 oracle on the
same fp16-rounded random-init weights.  Tolerances are the calibrated ones recorded in DESIGN.md §Parity:
   UNet eps      : max |err| <= 2e-2 * max|ref|   (fp16 activations, fp32 accumulate, ~60 fused layers)
   final latents : rel-L2 <= 1e-2
   frames        : mean |d| <= 1 LSB, >= 99 % of pixels within +-2 LSB, none beyond +-8 LSB
"""
import numpy as np
import pytest
import torch

from _helpers import MID_UNET, MID_VAE, TINY_UNET, TINY_VAE, make_oracle, product_cfgs

pytestmark = pytest.mark.gpu


def _engine(ocfg_u, ocfg_v, hw, frames, guidance=True, seed=0, tiled=False):
    from stable_diffusion_videos_b200.engine import Engine

    unet, vae = make_oracle(ocfg_u, ocfg_v, seed=seed)
    if tiled:
        from _helpers import set_tiled

        set_tiled(unet, vae)
    ucfg, vcfg = product_cfgs(ocfg_u, ocfg_v)
    eng = Engine(ucfg, vcfg, hw, frames, guidance=guidance, max_steps=64, tiled=tiled)
    eng.load_state_dict(unet.state_dict(), vae.state_dict())
    return eng, unet, vae


def _rel_l2(a, b):
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("cfgs,hw", [((TINY_UNET, TINY_VAE), (8, 8)), ((TINY_UNET, TINY_VAE), (16, 8)),
                                      ((MID_UNET, MID_VAE), (16, 16))])
def test_unet_forward(cfgs, hw):
    from stable_diffusion_videos_b200.schedulers import PNDMScheduler

    eng, unet, _ = _engine(cfgs[0], cfgs[1], hw, frames=1)
    sch = PNDMScheduler()
    eng.set_scheduler(sch, 4, 7.5)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, hw[0], hw[1], generator=g)
    ctx = torch.randn(2, 77, cfgs[0].cross_attention_dim, generator=g).half().float()
    for step in (0, 3):
        t = int(sch.timesteps[step])
        with torch.no_grad():
            ref = unet(x.half().float(), torch.tensor(t), ctx)
        out = eng.debug_unet(x.cuda(), step, ctx.cuda()).cpu()
        assert torch.isfinite(out).all()
        err = float((out - ref).abs().max())
        assert err <= 2e-2 * float(ref.abs().max()), (step, err, float(ref.abs().max()))
        assert _rel_l2(out, ref) <= 1e-2


@pytest.mark.parametrize("cfgs,hw,frames", [((TINY_UNET, TINY_VAE), (8, 8), 2), ((MID_UNET, MID_VAE), (16, 16), 1)])
def test_vae_decode(cfgs, hw, frames):
    eng, _, vae = _engine(cfgs[0], cfgs[1], hw, frames=frames)
    g = torch.Generator().manual_seed(5)
    lat = torch.randn(frames, 4, hw[0], hw[1], generator=g) * 0.18215 * 3
    with torch.no_grad():
        raw_ref = vae.decode(lat / 0.18215).permute(0, 2, 3, 1)
    u8, raw = eng.debug_vae(lat.cuda())
    raw = raw.cpu()
    assert torch.isfinite(raw).all()
    assert float((raw - raw_ref).abs().max()) <= 2e-2 * float(raw_ref.abs().max())
    ref_u8 = ((raw_ref / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    d = (u8.cpu().int() - ref_u8.int()).abs()
    assert d.float().mean() <= 1.0 and (d <= 2).float().mean() >= 0.99 and int(d.max()) <= 8


@pytest.mark.parametrize("kind,steps", [("pndm", 4), ("ddim", 5), ("lms", 6), ("euler", 5), ("dpm", 6)])
def test_full_sampler_tiny(kind, steps):
    from oracle.pipeline import sample_frames, synthetic_embedding, to_uint8
    from oracle.schedulers import make_scheduler
    from stable_diffusion_videos_b200.schedulers import SCHEDULERS

    hw, F = (8, 8), 2
    eng, unet, vae = _engine(TINY_UNET, TINY_VAE, hw, frames=F)
    D = TINY_UNET.cross_attention_dim
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(F, 4, *hw, generator=g)
    cond = torch.cat([synthetic_embedding(k, dim=D) for k in range(F)]).half().float()
    unc = synthetic_embedding("", dim=D).half().float()
    img, lat_ref, _ = sample_frames(unet, vae, make_scheduler(kind), lat, cond, unc, steps, 7.5, return_latents=True)
    eng.set_scheduler(SCHEDULERS[kind](), steps, 7.5)
    for use_graph in (False, True, True):
        u8, fin = eng.sample(lat.cuda(), cond.cuda(), unc.cuda(), use_graph=use_graph, return_latents=True)
        torch.cuda.synchronize()
        assert _rel_l2(fin.cpu(), lat_ref) <= 1e-2, (kind, use_graph, _rel_l2(fin.cpu(), lat_ref))
        d = np.abs(u8.cpu().numpy().astype(np.int32) - to_uint8(img).astype(np.int32))
        assert d.mean() <= 1.0 and (d <= 2).mean() >= 0.99 and d.max() <= 8, (kind, d.mean(), d.max())


@pytest.mark.parametrize("cfgs,hw,F", [((TINY_UNET, TINY_VAE), (8, 8), 2), ((MID_UNET, MID_VAE), (16, 16), 1)])
def test_tiled_circular_padding_full_sampler(cfgs, hw, F):
    """from_pretrained(tiled=True) (stable_diffusion_pipeline.py:841-858): every 3x3 conv of UNet and VAE pads circularly
    (stride-2 downsamplers, the folded nearest-up convs and the 4-channel edge convs included).  Also checks that the
    result really differs from the zero-padded one (the test would otherwise pass with the flag ignored)."""
    from oracle.pipeline import sample_frames, synthetic_embedding, to_uint8
    from oracle.schedulers import make_scheduler
    from stable_diffusion_videos_b200.schedulers import PNDMScheduler

    eng, unet, vae = _engine(cfgs[0], cfgs[1], hw, frames=F, tiled=True)
    D = cfgs[0].cross_attention_dim
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(F, 4, *hw, generator=g)
    cond = torch.cat([synthetic_embedding(k, dim=D) for k in range(F)]).half().float()
    unc = synthetic_embedding("", dim=D).half().float()
    img, lat_ref, raw_ref = sample_frames(unet, vae, make_scheduler("pndm"), lat, cond, unc, 3, 7.5, return_latents=True)
    eng.set_scheduler(PNDMScheduler(), 3, 7.5)
    u8, fin = eng.sample(lat.cuda(), cond.cuda(), unc.cuda(), use_graph=True, return_latents=True)
    torch.cuda.synchronize()
    assert _rel_l2(fin.cpu(), lat_ref) <= 1e-2, _rel_l2(fin.cpu(), lat_ref)
    d = np.abs(u8.cpu().numpy().astype(np.int32) - to_uint8(img).astype(np.int32))
    assert d.mean() <= 1.0 and (d <= 2).mean() >= 0.99 and d.max() <= 8, (d.mean(), d.max())
    eng0, _, _ = _engine(cfgs[0], cfgs[1], hw, frames=F, tiled=False)
    eng0.set_scheduler(PNDMScheduler(), 3, 7.5)
    _, fin0 = eng0.sample(lat.cuda(), cond.cuda(), unc.cuda(), use_graph=True, return_latents=True)
    assert _rel_l2(fin0.cpu(), lat_ref) > 3e-2  # zero padding is a different function


def test_slerp_lerp_batch_matches_reference_semantics():
    from oracle.slerp import slerp, slerp_f64
    from stable_diffusion_videos_b200 import _native as n

    for dtype, tol in ((torch.float32, 2e-6), (torch.float16, 1.0)):
        g = torch.Generator().manual_seed(11)
        la = torch.randn(1, 4, 64, 64, generator=g).to(dtype)
        lb = torch.randn(1, 4, 64, 64, generator=g).to(dtype)
        ea = torch.randn(1, 77, 768, generator=g).to(dtype)
        eb = torch.randn(1, 77, 768, generator=g).to(dtype)
        T = np.linspace(0.0, 1.0, 7)
        ol, oe = n.slerp_lerp_batch(la.cuda(), lb.cuda(), ea.cuda(), eb.cuda(), torch.tensor(T, dtype=torch.float32).cuda())
        torch.cuda.synchronize()
        ol, oe = ol.cpu(), oe.cpu()
        assert torch.equal(ol[0], la[0]) and torch.equal(ol[-1], lb[0])  # endpoints exact (utils.py semantics)
        for i, t in enumerate(T):
            truth = torch.from_numpy(slerp_f64(float(t), la.double().numpy(), lb.double().numpy()))[0]
            if dtype == torch.float32:
                assert float((ol[i].double() - truth).abs().max()) <= tol * 5
            else:  # fp32 math rounded once: <= 1 fp16 ulp of the fp64 truth + the fp32 rounding of the two products
                ulp = torch.finfo(torch.float16).eps * truth.abs().clamp_min(2.0 ** -14)
                slack = 2.0 ** -22 * (la[0].double().abs() + lb[0].double().abs())
                assert bool(((ol[i].double() - truth).abs() <= ulp + slack).all())
            ref_e = torch.lerp(ea.float(), eb.float(), float(t))[0]
            assert float((oe[i].float() - ref_e).abs().max()) <= (1e-6 if dtype == torch.float32 else 4e-3)


def test_slerp_colinear_falls_back_to_lerp():
    from stable_diffusion_videos_b200 import _native as n

    a = torch.randn(1, 4, 8, 8)
    b = a * 1.5
    e = torch.zeros(1, 8, 8)
    ol, _ = n.slerp_lerp_batch(a.cuda(), b.cuda(), e.cuda(), e.cuda(), torch.tensor([0.25]).cuda())
    assert torch.allclose(ol.cpu()[0], (0.75 * a + 0.25 * b)[0], atol=1e-6)


def test_missing_param_fails_loudly():
    from stable_diffusion_videos_b200 import _native as n
    from stable_diffusion_videos_b200.engine import Engine

    unet, vae = make_oracle(TINY_UNET, TINY_VAE)
    ucfg, vcfg = product_cfgs(TINY_UNET, TINY_VAE)
    eng = Engine(ucfg, vcfg, (8, 8), 1)
    sd = dict(unet.state_dict())
    sd.pop("conv_in.weight")
    with pytest.raises(n.SdwError):
        eng.load_state_dict(sd, vae.state_dict())
