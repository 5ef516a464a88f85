"""Native CLIP text tower against the INSTALLED transformers.CLIPTextModel (random-init, weights rounded to fp16, fp32
math on the CPU) — the one oracle of this repo that is third-party code, not a restatement (reference
stable_diffusion_pipeline.py:809-820, 341-348: `self.text_encoder(input_ids)[0]`).

Tolerance: fp16 activations through 12 (23) pre-LN layers against an fp32 run of the same fp16-rounded weights:
max |err| <= 2e-2 * max|ref| and rel-L2 <= 5e-3 (the UNet forward bound of tests/test_engine_gpu.py)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _hf(cfg_kw, seed=0):
    from transformers import CLIPTextConfig, CLIPTextModel

    torch.manual_seed(seed)
    model = CLIPTextModel(CLIPTextConfig(**cfg_kw)).eval()
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(p.half().float())
        # random-init LayerNorm affines / biases are 1 / 0: perturb them so that every parameter is exercised
        for n, p in model.named_parameters():
            if n.endswith("bias") or "layer_norm" in n:
                p.add_((torch.randn_like(p) * 0.05).half().float())
    return model


@pytest.mark.parametrize("name,cfg_kw,B", [
    ("small", dict(vocab_size=1000, hidden_size=256, intermediate_size=1024, num_hidden_layers=3, num_attention_heads=4,
                   max_position_embeddings=77, hidden_act="quick_gelu"), 3),
    ("sd1x-ViT-L", dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                        num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu"), 2),
    ("sd2x-OpenCLIP-H-23", dict(vocab_size=49408, hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                                num_attention_heads=16, max_position_embeddings=77, hidden_act="gelu"), 1),
])
def test_native_clip_matches_transformers(name, cfg_kw, B):
    from stable_diffusion_videos_b200.clip import NativeCLIPTextEncoder

    model = _hf(cfg_kw)
    g = torch.Generator().manual_seed(1)
    ids = torch.randint(0, cfg_kw["vocab_size"], (B, 77), generator=g)
    ids[:, 0] = cfg_kw["vocab_size"] - 2          # BOS-like
    ids[0, 20:] = cfg_kw["vocab_size"] - 1        # a padded prompt: EOS repeated to the end
    with torch.no_grad():
        ref = model(ids)[0]
    enc = NativeCLIPTextEncoder.from_hf_model(model, max_batch=2)
    out = enc(ids)[0]
    torch.cuda.synchronize()
    assert out.shape == ref.shape and out.dtype == torch.float16 and torch.isfinite(out).all()
    out = out.float().cpu()
    err = float((out - ref).abs().max())
    rel = float((out - ref).norm() / ref.norm())
    assert err <= 2e-2 * float(ref.abs().max()) and rel <= 5e-3, (name, err, float(ref.abs().max()), rel)


def test_native_clip_rejects_wrong_weights():
    from stable_diffusion_videos_b200 import _native
    from stable_diffusion_videos_b200.clip import NativeCLIPTextEncoder

    enc = NativeCLIPTextEncoder(vocab_size=100, hidden_size=128, num_hidden_layers=1, num_attention_heads=2,
                                intermediate_size=256)
    with pytest.raises(_native.SdwError):
        enc.load_state_dict({"text_model.final_layer_norm.weight": torch.zeros(64)})
    with pytest.raises(_native.SdwError):
        enc(torch.zeros(1, 77, dtype=torch.long))  # parameters missing: the forward refuses
