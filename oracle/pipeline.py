"""ORACLE (test infrastructure only).  CPU restatement of the reference's hot-path control flow:

  generate_inputs   stable_diffusion_pipeline.py:457-479   (lerp embeddings P:467, slerp noise P:468)
  __call__          stable_diffusion_pipeline.py:310-358 (CFG batch), 365-401 (latents / scheduler set-up),
                    412-430 (denoise loop), 432-438 (decode + post-process), 449-450 (numpy_to_pil quantisation:
                    `(x * 255).round().astype("uint8")`, a diffusers helper)
  init_noise        stable_diffusion_pipeline.py:822-838

The text encoder is a *feed* of the hot path, not part of it: the oracle takes precomputed embeddings
(`embed_fn(prompt) -> [1, 77, D]`), which is also how BASELINE.json's configs are stated (synthetic prompt
embeddings).  PARITY UNPINNED (see oracle/__init__.py).
"""
import numpy as np
import torch

from .slerp import slerp


def synthetic_embedding(key, tokens=77, dim=768, dtype=torch.float32):
    """cond embeddings `randn([1,77,D], Generator(cpu).manual_seed(1000 + k))`, uncond seed 999."""
    seed = 999 if key in ("", None) else 1000 + int(key)
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn((1, tokens, dim), generator=g, dtype=torch.float32).to(dtype)


def init_noise(seed, noise_shape, dtype, device="cpu"):
    """stable_diffusion_pipeline.py:832-837 — seeded torch.randn on the pipeline device."""
    return torch.randn(noise_shape, device=device, generator=torch.Generator(device=device).manual_seed(seed),
                       dtype=dtype)


def generate_inputs(embeds_a, embeds_b, latents_a, latents_b, T, batch_size):
    """stable_diffusion_pipeline.py:464-479 with the embeddings / keyframe noise already computed."""
    batch_idx = 0
    eb, nb = None, None
    for i, t in enumerate(T):
        embeds = torch.lerp(embeds_a, embeds_b, float(t))
        noise = slerp(float(t), latents_a, latents_b)
        eb = embeds if eb is None else torch.cat([eb, embeds])
        nb = noise if nb is None else torch.cat([nb, noise])
        if not (eb.shape[0] == batch_size or i + 1 == T.shape[0]):
            continue
        yield batch_idx, eb, nb
        batch_idx += 1
        eb, nb = None, None


@torch.no_grad()
def sample_frames(unet, vae, scheduler, latents, text_embeddings, uncond_embeddings, num_inference_steps=50,
                  guidance_scale=7.5, eta=0.0, return_latents=False, callback=None):
    """stable_diffusion_pipeline.py:310-438 given interpolated (latents, text_embeddings) batches.

    Returns float32 NHWC images in [0, 1] (the array the reference hands to numpy_to_pil) and, optionally, the
    final latents and the pre-clamp decoder output (for saturation-proof parity checks).
    """
    B = text_embeddings.shape[0]
    do_cfg = guidance_scale > 1.0
    if do_cfg:
        unc = uncond_embeddings.repeat(B, 1, 1)
        text_embeddings = torch.cat([unc, text_embeddings])  # P:352-358
    scheduler.set_timesteps(num_inference_steps)  # P:394
    latents = latents * scheduler.init_noise_sigma  # P:401
    for i, t in enumerate(scheduler.timesteps):
        x = torch.cat([latents] * 2) if do_cfg else latents  # P:414
        x = scheduler.scale_model_input(x, t)  # P:415
        noise_pred = unet(x, t, text_embeddings)  # P:418
        if do_cfg:
            u, c = noise_pred.chunk(2)
            noise_pred = u + guidance_scale * (c - u)  # P:423
        latents = scheduler.step(noise_pred, t, latents)  # P:426
        if callback is not None:
            callback(i, t, latents)
    final_latents = latents
    image_raw = vae.decode(1 / 0.18215 * latents)  # P:432-433
    image = (image_raw / 2 + 0.5).clamp(0, 1)  # P:435
    image = image.cpu().permute(0, 2, 3, 1).float().numpy()  # P:438
    if return_latents:
        return image, final_latents, image_raw
    return image


def to_uint8(images):
    """diffusers `numpy_to_pil` quantisation used at stable_diffusion_pipeline.py:450."""
    return (images * 255).round().astype("uint8")


def walk_frames(unet, vae, scheduler, keys, seeds, num_interpolation_steps, latent_hw, batch_size=1,
                num_inference_steps=50, guidance_scale=7.5, embed_dim=768, T=None):
    """walk() (stable_diffusion_pipeline.py:731-785) reduced to the frame-producing core: for each consecutive
    (key, seed) pair, `num_interpolation_steps` frames; returns uint8 [n, H, W, 3]."""
    frames = []
    uncond = synthetic_embedding("", dim=embed_dim)
    h, w = latent_hw
    for k in range(len(keys) - 1):
        ea, eb = synthetic_embedding(keys[k], dim=embed_dim), synthetic_embedding(keys[k + 1], dim=embed_dim)
        la = init_noise(seeds[k], (1, unet.cfg.in_channels, h, w), ea.dtype)
        lb = init_noise(seeds[k + 1], (1, unet.cfg.in_channels, h, w), ea.dtype)
        n = num_interpolation_steps[k] if isinstance(num_interpolation_steps, (list, tuple)) else num_interpolation_steps
        Tk = np.linspace(0.0, 1.0, n) if T is None else T  # P:509
        for _, e, z in generate_inputs(ea, eb, la, lb, Tk, batch_size):
            frames.append(to_uint8(sample_frames(unet, vae, scheduler, z, e, uncond, num_inference_steps,
                                                 guidance_scale)))
    return np.concatenate(frames, axis=0)
