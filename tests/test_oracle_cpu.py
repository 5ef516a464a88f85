"""CPU tests pinning the oracle (not gpu-marked): structural KATs, scheduler tables, slerp identities, and the
host-side scheduler plans of the product against the oracle's stateful schedulers."""
import numpy as np
import pytest
import torch

from _helpers import OUNetConfig, OVAEConfig, TINY_UNET, TINY_VAE, make_oracle, product_cfgs
from oracle import schedulers as O
from oracle.slerp import slerp, slerp_f64
from oracle.unet import UNet2DConditionModel
from oracle.vae import AutoencoderKLDecoder


def test_published_parameter_counts():
    with torch.device("meta"):
        assert sum(p.numel() for p in UNet2DConditionModel(OUNetConfig.sd14()).parameters()) == 859_520_964
        assert sum(p.numel() for p in UNet2DConditionModel(OUNetConfig.sd21()).parameters()) == 865_910_724
        assert sum(p.numel() for p in AutoencoderKLDecoder(OVAEConfig()).parameters()) == 49_490_199


def test_product_shape_tables_match_oracle_state_dicts():
    from stable_diffusion_videos_b200.configs import UNetConfig, VAEConfig, unet_param_shapes, vae_param_shapes

    with torch.device("meta"):
        for oc, pc in ((OUNetConfig.sd14(), UNetConfig.sd14()), (OUNetConfig.sd21(), UNetConfig.sd21())):
            sd = {k: tuple(v.shape) for k, v in UNet2DConditionModel(oc).state_dict().items()}
            assert sd == unet_param_shapes(pc)
        sd = {k: tuple(v.shape) for k, v in AutoencoderKLDecoder(OVAEConfig()).state_dict().items()}
        assert sd == vae_param_shapes(VAEConfig())


def test_scheduler_timestep_kats():
    s = O.PNDMScheduler()
    s.set_timesteps(50)
    ts = s.timesteps.tolist()
    assert len(ts) == 51 and ts[:4] == [981, 961, 961, 941] and ts[-2:] == [21, 1]
    s.set_timesteps(4)
    assert s.timesteps.tolist() == [751, 501, 501, 251, 1]
    d = O.DDIMScheduler()
    d.set_timesteps(50)
    assert d.timesteps.tolist()[:2] == [981, 961] and d.timesteps.tolist()[-1] == 1 and len(d.timesteps) == 50
    l = O.LMSDiscreteScheduler()
    l.set_timesteps(50)
    assert abs(l.init_noise_sigma - 14.6146) < 1e-3 and l.timesteps[0] == 999.0 and l.timesteps[-1] == 0.0


def _replay(plan, x0, eps_list, sigma):
    x, hist, xb = x0 * sigma, [None] * 4, None
    for st, e in zip(plan, eps_list):
        s = x
        if st["save_x_base"]:
            xb = x
        if st["use_x_base"]:
            s = xb
        acc = st["c_e"][0] * e
        for j in range(4):
            if st["c_e"][j + 1] != 0:
                acc = acc + st["c_e"][j + 1] * hist[st["hist_slot"][j]]
        if st["push_slot"] >= 0:  # what sdw_cfg_sched_step keeps: push_e * e + push_x * s (eps history by default)
            hist[st["push_slot"]] = st.get("push_e", 1.0) * e + st.get("push_x", 0.0) * s
        x = st["c_x"] * s + acc
    return x


@pytest.mark.parametrize("kind,n,pt", [("pndm", 50, "epsilon"), ("pndm", 4, "epsilon"), ("pndm", 1, "epsilon"),
                                        ("ddim", 50, "epsilon"), ("ddim", 50, "v_prediction"),
                                        ("lms", 50, "epsilon"), ("lms", 3, "epsilon"), ("euler", 50, "epsilon"),
                                        ("euler", 7, "v_prediction"), ("dpm", 50, "epsilon"), ("dpm", 25, "epsilon"),
                                        ("dpm", 10, "epsilon"), ("dpm", 1, "epsilon")])
def test_product_scheduler_plan_equals_oracle_scheduler(kind, n, pt):
    from stable_diffusion_videos_b200.schedulers import SCHEDULERS

    p = SCHEDULERS[kind](prediction_type=pt)
    p.set_timesteps(n)
    plan = p.plan()
    o = O.make_scheduler(kind, pt)
    o.set_timesteps(n)
    assert np.allclose(np.asarray(p.timesteps, dtype=float), o.timesteps.numpy().astype(float))
    g = torch.Generator().manual_seed(0)
    x0 = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64)
    eps = [torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) for _ in plan]
    xo = x0 * o.init_noise_sigma
    for t, e in zip(o.timesteps, eps):
        xo = o.step(e, t, xo)
    xn = _replay(plan, x0, eps, p.init_noise_sigma)
    assert float((xo - xn).abs().max()) <= 1e-6 * float(xo.abs().max()) + 1e-9
    if kind in ("lms", "euler"):
        for i, t in enumerate(o.timesteps):
            assert abs(float(o.scale_model_input(torch.ones(1), t)) - plan[i]["in_scale"]) < 1e-6


def test_slerp_kats():
    g = lambda s: torch.randn((1, 4, 64, 64), generator=torch.Generator("cpu").manual_seed(s))
    a, b = g(42), g(1337)
    an, bn = a.numpy(), b.numpy()
    dot = float(np.sum(an * bn / (np.linalg.norm(an) * np.linalg.norm(bn))))
    assert abs(dot - (-0.00170)) < 2e-4
    assert torch.equal(slerp(0.0, a, b), a) and torch.equal(slerp(1.0, a, b), b)
    for t in (0.1, 0.5, 0.9):
        truth = slerp_f64(t, an, bn)
        assert np.abs(slerp(t, a, b).numpy() - truth).max() < 2e-6
        h = slerp(t, a.half(), b.half()).float().numpy()
        assert np.abs(h - truth).max() < 4e-3
    # colinear inputs take the lerp branch (utils.py:51-52)
    assert torch.allclose(slerp(0.25, a, a * 2.0), 0.75 * a + 0.25 * (a * 2.0))


def test_oracle_tiny_pipeline_is_informative():
    """sanity gate of SURVEY.md §8d: random-init frames must be neither saturated nor NaN."""
    from oracle.pipeline import walk_frames

    unet, vae = make_oracle(TINY_UNET, TINY_VAE)
    frames = walk_frames(unet, vae, O.make_scheduler("pndm"), [0, 1], [42, 1337], 2, (8, 8), batch_size=2,
                         num_inference_steps=4, embed_dim=TINY_UNET.cross_attention_dim)
    assert frames.shape == (2, 16, 16, 3) and frames.dtype == np.uint8
    sat = ((frames == 0) | (frames == 255)).mean()
    assert sat < 0.2
