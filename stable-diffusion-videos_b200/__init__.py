"""stable-diffusion-videos_b200 — Blackwell-native latent-walk hot path.

Drop-in surface mirrors `stable_diffusion_videos` (reference __init__.py:99-119) for the hot path:
StableDiffusionWalkPipeline (+ slerp).  Import as `stable_diffusion_videos_b200`.
"""
__version__ = "0.1.0"
