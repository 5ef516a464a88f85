"""Golden fixtures at the BENCHMARKED configuration (BASELINE.json configs[1]: SD-1.4 full widths, 64x64 latents ->
512x512 frames), generated with the ORACLE (oracle/ — CPU fp32 restatement; the reference cannot run here), plus the
tolerance CALIBRATION SURVEY.md §8d asks for: the same run with fp16-storage emulation (every module output, the
latents and the scheduler state rounded to fp16 — what the reference's fp16 CUDA pipeline stores), whose spread
against the fp32 oracle is the yardstick for the native path.  Run from the repo root (~15 min on 8 cores):

    python tests/golden/make_golden_full.py [case ...]

Cases (weights: oracle modules under torch.manual_seed(0), fp16-rounded; inputs as SURVEY.md §8d: embeddings
seeds 1000/1001, uncond 999, init_noise seeds 42/1337, guidance 7.5):
  full_pndm10_f2 : F = 2 frames (T = [0, 0.5] of the 2-prompt walk), PNDM 10 steps (11 UNet calls, batch 4)
  full_pndm50_f1 : F = 1 frame  (T = [0.5]),                          PNDM 50 steps (51 UNet calls, batch 2)
  full_lms50_f1  : F = 1 frame  (T = [0.5]), K-LMS 50 steps, guidance 15  (BASELINE configs[4] sampler, EX-MV:15-17,43-54)
  sd21_ddim50_f1 : SD-2.1 structure (heads (5,10,20,20) => d = 64, ctx 1024, linear projections, v-prediction),
                   widths (320,640,1280,1280), 24x24 latents, DDIM 50 steps, F = 1  (BASELINE configs[3] numerics)

Stored per case in <case>.npz: final latents fp32, uint8 frames, pre-clamp decoder output subsampled (stride 4, fp16);
calibration numbers in calibration.json.
"""
import json
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

RAW_STRIDE = 4

CASES = {
    "full_pndm10_f2": dict(model="sd14", hw=(64, 64), T=[0.0, 0.5], sched="pndm", steps=10, guidance=7.5),
    "full_pndm50_f1": dict(model="sd14", hw=(64, 64), T=[0.5], sched="pndm", steps=50, guidance=7.5),
    "full_lms50_f1": dict(model="sd14", hw=(64, 64), T=[0.5], sched="lms", steps=50, guidance=15.0),
    "sd21_ddim50_f1": dict(model="sd21", hw=(24, 24), T=[0.5], sched="ddim_v", steps=50, guidance=7.5),
}


def model_cfgs(model):
    if model == "sd14":
        return OUNetConfig.sd14(), OVAEConfig()
    return OUNetConfig.sd21(), OVAEConfig()


def scheduler_for(kind):
    if kind == "ddim_v":
        return make_scheduler("ddim", prediction_type="v_prediction")
    return make_scheduler(kind)


def case_inputs(case):
    c = CASES[case]
    ucfg, _ = model_cfgs(c["model"])
    D = ucfg.cross_attention_dim
    ea, eb = synthetic_embedding(0, dim=D).half().float(), synthetic_embedding(1, dim=D).half().float()
    unc = synthetic_embedding("", dim=D).half().float()
    la, lb = init_noise(42, (1, 4, *c["hw"]), torch.float32), init_noise(1337, (1, 4, *c["hw"]), torch.float32)
    T = np.asarray(c["T"], dtype=np.float64)
    (_, emb, lat), = list(generate_inputs(ea, eb, la, lb, T, len(T)))
    return dict(ea=ea, eb=eb, unc=unc, la=la, lb=lb, T=T, emb=emb, lat=lat)


class Fp16Storage:
    """fp16-storage emulation: every module output is rounded to fp16 (math stays fp32 = fp32 accumulate)."""

    def __init__(self, *modules):
        self.handles = []
        for m in modules:
            for sub in m.modules():
                self.handles.append(sub.register_forward_hook(self._hook))

    @staticmethod
    def _hook(_mod, _inp, out):
        if isinstance(out, torch.Tensor) and out.is_floating_point():
            return out.half().float()
        return out

    def remove(self):
        for h in self.handles:
            h.remove()


def run_case(case, unet, vae, fp16_storage=False):
    c, inp = CASES[case], case_inputs(case)
    cb = None
    emu = None
    lat = inp["lat"]
    if fp16_storage:
        emu = Fp16Storage(unet, vae)
        lat = lat.half().float()

        def cb(_i, _t, latents):  # the reference's latents live in fp16 between steps
            latents.copy_(latents.half().float())
    try:
        img, fin, raw = sample_frames(unet, vae, scheduler_for(c["sched"]), lat, inp["emb"], inp["unc"], c["steps"],
                                      c["guidance"], return_latents=True, callback=cb)
    finally:
        if emu:
            emu.remove()
    return to_uint8(img), fin.numpy().astype(np.float32), raw.permute(0, 2, 3, 1).numpy()


def spread(a, b):
    fa, la, ra = a
    fb, lb, rb = b
    d = np.abs(fa.astype(np.int32) - fb.astype(np.int32))
    return {
        "latents_rel_l2": float(np.linalg.norm(la - lb) / np.linalg.norm(lb)),
        "latents_max_abs_over_max": float(np.abs(la - lb).max() / np.abs(lb).max()),
        "raw_max_abs_over_max": float(np.abs(ra - rb).max() / np.abs(rb).max()),
        "raw_rel_l2": float(np.linalg.norm(ra - rb) / np.linalg.norm(rb)),
        "frames_mean_lsb": float(d.mean()), "frames_frac_within_2": float((d <= 2).mean()),
        "frames_p999_lsb": float(np.quantile(d, 0.999)), "frames_max_lsb": int(d.max()),
    }


def main(cases):
    cal_path = os.path.join(HERE, "calibration.json")
    cal = json.load(open(cal_path)) if os.path.exists(cal_path) else {}
    built = {}
    for case in cases:
        model = CASES[case]["model"]
        if model not in built:
            built.clear()
            built[model] = make_oracle(*model_cfgs(model), seed=0)
        unet, vae = built[model]
        t0 = time.time()
        ref = run_case(case, unet, vae)
        t1 = time.time()
        emu = run_case(case, unet, vae, fp16_storage=True)
        frames, fin, raw = ref
        sat = float(np.mean((frames == 0) | (frames == 255)))
        cal[case] = dict(spread(emu, ref), oracle_seconds=round(t1 - t0, 1), saturated_frac=sat,
                         raw_range=[float(raw.min()), float(raw.max())], latents_absmax=float(np.abs(fin).max()),
                         note="fp16-storage-emulated oracle vs fp32 oracle, same fp16-rounded weights")
        np.savez_compressed(os.path.join(HERE, case + ".npz"), frames=frames, latents=fin,
                            raw=raw[:, ::RAW_STRIDE, ::RAW_STRIDE].astype(np.float16))
        print(case, json.dumps(cal[case]), flush=True)
        json.dump(cal, open(cal_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
