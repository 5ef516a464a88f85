"""Generates the golden fixtures of BASELINE.json configs[0] ("2-prompt walk, 2 interp steps, 64x64, 4 inference
steps, random-init SD-1.4 on torch CPU") with the ORACLE (oracle/ — the CPU restatement of the reference path; the
reference itself cannot run here: diffusers is not installable).  Run from the repo root:

    python tests/golden/make_golden.py

Weights: oracle modules built under torch.manual_seed(0) (UNet first, then VAE decoder), rounded to fp16.
Inputs: synthetic embeddings keys 0/1 (seeds 1000/1001), uncond seed 999, init_noise seeds 42/1337 (CPU generator),
T = linspace(0, 1, 2), guidance 7.5, PNDM 4 steps (5 UNet calls), 8x8 latents -> 64x64 frames.
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))

from _helpers import OUNetConfig, OVAEConfig, make_oracle  # noqa: E402
from oracle.pipeline import generate_inputs, init_noise, sample_frames, synthetic_embedding, to_uint8  # noqa: E402
from oracle.schedulers import make_scheduler  # noqa: E402


def run():
    t0 = time.time()
    unet, vae = make_oracle(OUNetConfig.sd14(), OVAEConfig(), seed=0)
    ea, eb = synthetic_embedding(0).half().float(), synthetic_embedding(1).half().float()
    unc = synthetic_embedding("").half().float()
    la, lb = init_noise(42, (1, 4, 8, 8), torch.float32), init_noise(1337, (1, 4, 8, 8), torch.float32)
    (_, emb, lat), = list(generate_inputs(ea, eb, la, lb, np.linspace(0.0, 1.0, 2), 2))
    img, fin, raw = sample_frames(unet, vae, make_scheduler("pndm"), lat, emb, unc, 4, 7.5, return_latents=True)
    print(f"oracle cfg-1: {time.time() - t0:.1f}s; raw range [{raw.min():.3f}, {raw.max():.3f}], "
          f"saturated {np.mean((to_uint8(img) == 0) | (to_uint8(img) == 255)):.3%}")
    return to_uint8(img), fin.numpy(), raw.permute(0, 2, 3, 1).numpy()


if __name__ == "__main__":
    frames, latents, raw = run()
    np.save(os.path.join(HERE, "cfg1_frames_u8.npy"), frames)
    np.save(os.path.join(HERE, "cfg1_final_latents.npy"), latents.astype(np.float32))
    np.save(os.path.join(HERE, "cfg1_raw_f16.npy"), raw.astype(np.float16))
    print("wrote", frames.shape, latents.shape)
