"""Fabricate a tiny LOCAL diffusers-layout checkpoint directory (unet/, vae/, text_encoder/, tokenizer/, scheduler/) from
the oracle's random-init modules — what `StableDiffusionWalkPipeline.from_pretrained` (reference P:841-858) reads.
The VAE attention keys are written with the OLD diffusers names (query / key / value / proj_attn) so that the loader's
aliasing is exercised."""
import json
import os

import torch

from _helpers import TINY_UNET, TINY_VAE, make_oracle

OLD_VAE_NAMES = {"to_q": "query", "to_k": "key", "to_v": "value", "to_out.0": "proj_attn"}


def _old_vae_key(k):
    if ".attentions." not in k:
        return k
    for new, old in OLD_VAE_NAMES.items():
        k = k.replace("." + new + ".", "." + old + ".")
    return k


def write_tokenizer(path):
    os.makedirs(path, exist_ok=True)
    chars = list("abcdefghijklmnopqrstuvwxyz0123456789")
    toks = chars + [c + "</w>" for c in chars] + ["<|startoftext|>", "<|endoftext|>"]
    vocab = {t: i for i, t in enumerate(toks)}
    json.dump(vocab, open(os.path.join(path, "vocab.json"), "w"))
    open(os.path.join(path, "merges.txt"), "w").write("#version: 0.2\n")
    json.dump({"model_max_length": 77, "bos_token": "<|startoftext|>", "eos_token": "<|endoftext|>",
               "unk_token": "<|endoftext|>", "pad_token": "<|endoftext|>", "tokenizer_class": "CLIPTokenizer"},
              open(os.path.join(path, "tokenizer_config.json"), "w"))
    return len(toks)


def write_checkpoint(root, scheduler="PNDMScheduler", scheduler_extra=None, with_weights=True):
    from safetensors.torch import save_file

    os.makedirs(root, exist_ok=True)
    u, v = TINY_UNET, TINY_VAE
    for sub in ("unet", "vae", "scheduler"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    json.dump({"_class_name": "UNet2DConditionModel", "in_channels": u.in_channels, "out_channels": u.out_channels,
               "block_out_channels": list(u.block_out_channels), "layers_per_block": u.layers_per_block,
               "attention_head_dim": u.attention_head_dim, "cross_attention_dim": u.cross_attention_dim,
               "norm_num_groups": u.norm_num_groups, "norm_eps": u.norm_eps, "sample_size": u.sample_size,
               "use_linear_projection": u.use_linear_projection}, open(os.path.join(root, "unet", "config.json"), "w"))
    json.dump({"_class_name": "AutoencoderKL", "latent_channels": v.latent_channels, "out_channels": v.out_channels,
               "block_out_channels": list(v.block_out_channels), "layers_per_block": v.layers_per_block,
               "norm_num_groups": v.norm_num_groups}, open(os.path.join(root, "vae", "config.json"), "w"))
    sc = {"_class_name": scheduler, "num_train_timesteps": 1000, "beta_start": 0.00085, "beta_end": 0.012,
          "beta_schedule": "scaled_linear", "skip_prk_steps": True, "set_alpha_to_one": False, "steps_offset": 1,
          "trained_betas": None}
    sc.update(scheduler_extra or {})
    json.dump(sc, open(os.path.join(root, "scheduler", "scheduler_config.json"), "w"))
    vocab = write_tokenizer(os.path.join(root, "tokenizer"))
    if not with_weights:
        return None
    unet, vae = make_oracle(u, v)
    save_file({k: t.half().contiguous() for k, t in unet.state_dict().items()},
              os.path.join(root, "unet", "diffusion_pytorch_model.safetensors"))
    save_file({_old_vae_key(k): (t.half().reshape(t.shape[0], t.shape[1]) if t.dim() == 4 and ".attentions." in k else t.half()).contiguous()
               for k, t in vae.state_dict().items()},
              os.path.join(root, "vae", "diffusion_pytorch_model.safetensors"))
    from transformers import CLIPTextConfig, CLIPTextModel

    torch.manual_seed(3)
    te = CLIPTextModel(CLIPTextConfig(vocab_size=vocab, hidden_size=u.cross_attention_dim, intermediate_size=128,
                                      num_hidden_layers=2, num_attention_heads=1, max_position_embeddings=77,
                                      hidden_act="quick_gelu", bos_token_id=vocab - 2, eos_token_id=vocab - 1,
                                      pad_token_id=vocab - 1)).eval()
    te.save_pretrained(os.path.join(root, "text_encoder"))
    return unet, vae, te
