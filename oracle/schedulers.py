"""ORACLE (test infrastructure only).  Stateful restatement of the three `diffusers` schedulers the reference
drives through `set_timesteps / scale_model_input / step / init_noise_sigma`
(stable_diffusion_pipeline.py:394, 401, 415, 426): PNDMScheduler (skip_prk_steps=True — the SD-1.4 default),
DDIMScheduler (SD-2.1 default, epsilon or v-prediction, eta = 0) and LMSDiscreteScheduler
(examples/make_music_video.py:15-17).  diffusers is un-vendored / un-pinned (pyproject.toml:14); the update
rules below restate its published v0.11 code paths.  The reference forces steps_offset=1 and
clip_sample=False (stable_diffusion_pipeline.py:85-110).  PARITY UNPINNED — pinned only by the timestep-table
KATs in tests/test_oracle_cpu.py.
"""
import numpy as np
import torch
from scipy import integrate


def _alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000):
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2  # "scaled_linear"
    return torch.cumprod(1.0 - betas, dim=0)


class PNDMScheduler:
    """PLMS variant (skip_prk_steps=True), steps_offset=1, set_alpha_to_one=False."""

    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, prediction_type="epsilon"):
        self.n_train = num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(n=num_train_timesteps)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.prediction_type = prediction_type
        self.timesteps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        r = self.n_train // n
        ts = (np.arange(0, n) * r).round() + 1
        self.timesteps = torch.from_numpy(
            np.concatenate([ts[:-1], ts[-2:-1], ts[-1:]])[::-1].copy().astype(np.int64))
        self.ets = []
        self.counter = 0
        self.cur_sample = None

    def scale_model_input(self, sample, t):
        return sample

    def step(self, model_output, timestep, sample):
        timestep = int(timestep)
        r = self.n_train // self.num_inference_steps
        prev = timestep - r
        if self.counter != 1:
            self.ets = self.ets[-3:]
            self.ets.append(model_output)
        else:
            prev = timestep
            timestep = timestep + r
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] + 37 * self.ets[-3] - 9 * self.ets[-4])
        self.counter += 1
        return self._prev_sample(sample, timestep, prev, model_output)

    def _prev_sample(self, sample, t, prev, model_output):
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t, b_p = 1 - a_t, 1 - a_p
        if self.prediction_type == "v_prediction":
            model_output = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        sample_coeff = (a_p / a_t) ** 0.5
        denom = a_t * b_p ** 0.5 + (a_t * b_t * a_p) ** 0.5
        return sample_coeff * sample - (a_p - a_t) * model_output / denom


class DDIMScheduler:
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, prediction_type="epsilon"):
        self.n_train = num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(n=num_train_timesteps)
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.prediction_type = prediction_type

    def set_timesteps(self, n):
        self.num_inference_steps = n
        r = self.n_train // n
        self.timesteps = torch.from_numpy(((np.arange(0, n) * r).round()[::-1].copy() + 1).astype(np.int64))

    def scale_model_input(self, sample, t):
        return sample

    def step(self, model_output, timestep, sample, eta=0.0):
        assert eta == 0.0, "oracle restates the deterministic (eta = 0) DDIM path only"
        timestep = int(timestep)
        prev = timestep - self.n_train // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        else:
            x0 = a_t ** 0.5 * sample - b_t ** 0.5 * model_output
            model_output = a_t ** 0.5 * model_output + b_t ** 0.5 * sample
        return a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * model_output


class LMSDiscreteScheduler:
    def __init__(self, num_train_timesteps=1000, prediction_type="epsilon"):
        self.n_train = num_train_timesteps
        self.alphas_cumprod = _alphas_cumprod(n=num_train_timesteps)
        ac = self.alphas_cumprod.numpy()
        sig = ((1 - ac) / ac) ** 0.5
        self.init_noise_sigma = float(np.concatenate([sig[::-1], [0.0]]).astype(np.float32).max())
        self.prediction_type = prediction_type

    def set_timesteps(self, n):
        self.num_inference_steps = n
        ts = np.linspace(0, self.n_train - 1, n, dtype=float)[::-1].copy()
        ac = self.alphas_cumprod.numpy()
        sig = np.array(((1 - ac) / ac) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = torch.from_numpy(ts)
        self.derivatives = []
        self._i = 0

    def _index(self, t):
        return int((self.timesteps == float(t)).nonzero()[0].item())

    def scale_model_input(self, sample, t):
        s = float(self.sigmas[self._index(t)])
        return sample / ((s ** 2 + 1) ** 0.5)

    def lms_coefficient(self, order, t, current_order):
        def f(tau):
            prod = 1.0
            for k in range(order):
                if current_order == k:
                    continue
                prod *= (tau - self.sigmas[t - k]) / (self.sigmas[t - current_order] - self.sigmas[t - k])
            return prod

        return integrate.quad(f, self.sigmas[t], self.sigmas[t + 1], epsrel=1e-4)[0]

    def step(self, model_output, timestep, sample, order=4):
        i = self._index(timestep)
        s = float(self.sigmas[i])
        if self.prediction_type == "epsilon":
            x0 = sample - s * model_output
        else:
            x0 = model_output * (-s / (s ** 2 + 1) ** 0.5) + (sample / (s ** 2 + 1))
        self.derivatives.append((sample - x0) / s)
        if len(self.derivatives) > order:
            self.derivatives.pop(0)
        order = min(i + 1, order)
        coeffs = [self.lms_coefficient(order, i, k) for k in range(order)]
        return sample + sum(c * d for c, d in zip(coeffs, reversed(self.derivatives)))


class EulerDiscreteScheduler:
    """diffusers EulerDiscreteScheduler, deterministic settings (s_churn = 0, timestep_spacing "linspace", linear sigma
    interpolation): the published Euler sampler of Karras et al. 2022 (Alg. 2).  diffusers is not installed; restated
    from the paper and the scheduler's documented behaviour, step by step (not via the coefficient form the product uses)."""

    def __init__(self, num_train_timesteps=1000, prediction_type="epsilon"):
        self.num_train_timesteps = num_train_timesteps
        self.prediction_type = prediction_type
        ac = _alphas_cumprod(n=num_train_timesteps).numpy()
        self._sig_all = np.array(((1 - ac) / ac) ** 0.5)
        self.init_noise_sigma = float(np.concatenate([self._sig_all[::-1], [0.0]]).astype(np.float32).max())

    def set_timesteps(self, n):
        ts = np.linspace(0, self.num_train_timesteps - 1, n, dtype=float)[::-1].copy()
        sig = np.interp(ts, np.arange(0, len(self._sig_all)), self._sig_all)
        self.sigmas = np.concatenate([sig, [0.0]]).astype(np.float32)
        self.timesteps = torch.from_numpy(ts)

    def _index(self, t):
        return int((self.timesteps == float(t)).nonzero()[0].item())

    def scale_model_input(self, sample, t):
        s = float(self.sigmas[self._index(t)])
        return sample / ((s ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample):
        i = self._index(timestep)
        s = float(self.sigmas[i])
        if self.prediction_type == "epsilon":
            x0 = sample - s * model_output
        else:
            x0 = model_output * (-s / (s ** 2 + 1) ** 0.5) + (sample / (s ** 2 + 1))
        derivative = (sample - x0) / s
        return sample + derivative * (float(self.sigmas[i + 1]) - s)


class DPMSolverMultistepScheduler:
    """diffusers v0.11 `DPMSolverMultistepScheduler` defaults (dpmsolver++, solver_order 2, midpoint, lower_order_final,
    epsilon prediction, no thresholding) — a member of the reference's scheduler union (stable_diffusion_pipeline.py:71-78).
    Stateful restatement of its `step`: convert_model_output, first-order update, multistep second-order update."""

    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, prediction_type="epsilon", solver_order=2, lower_order_final=True):
        assert prediction_type == "epsilon"
        self.n_train = num_train_timesteps
        ac = _alphas_cumprod(n=num_train_timesteps).double()
        self.alpha_t, self.sigma_t = ac.sqrt(), (1 - ac).sqrt()
        self.lambda_t = self.alpha_t.log() - self.sigma_t.log()
        self.solver_order, self.lower_order_final = solver_order, lower_order_final
        self.timesteps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        self.timesteps = torch.from_numpy(np.linspace(0, self.n_train - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64))
        self.model_outputs = [None] * self.solver_order
        self.lower_order_nums = 0
        self.prev_t = None

    def scale_model_input(self, sample, t):
        return sample

    def step(self, model_output, timestep, sample):
        t = int(timestep)
        idx = int((self.timesteps == t).nonzero()[0])
        prev = 0 if idx == len(self.timesteps) - 1 else int(self.timesteps[idx + 1])
        final = idx == len(self.timesteps) - 1 and self.lower_order_final and len(self.timesteps) < 15
        x0 = (sample - float(self.sigma_t[t]) * model_output) / float(self.alpha_t[t])  # convert_model_output
        self.model_outputs = self.model_outputs[1:] + [x0]
        lam_t, lam_s = float(self.lambda_t[prev]), float(self.lambda_t[t])
        h = lam_t - lam_s
        a_p, s_p, s_t = float(self.alpha_t[prev]), float(self.sigma_t[prev]), float(self.sigma_t[t])
        if self.solver_order == 1 or self.lower_order_nums < 1 or final:
            out = (s_p / s_t) * sample - a_p * np.expm1(-h) * x0
        else:
            m0, m1 = self.model_outputs[-1], self.model_outputs[-2]
            h0 = lam_s - float(self.lambda_t[self.prev_t])
            r0 = h0 / h
            d1 = (1.0 / r0) * (m0 - m1)
            out = (s_p / s_t) * sample - a_p * np.expm1(-h) * m0 - 0.5 * a_p * np.expm1(-h) * d1
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.prev_t = t
        return out.to(sample.dtype)


def make_scheduler(kind, prediction_type="epsilon"):
    return {"dpm": DPMSolverMultistepScheduler, "pndm": PNDMScheduler, "ddim": DDIMScheduler, "lms": LMSDiscreteScheduler,
            "euler": EulerDiscreteScheduler}[kind](prediction_type=prediction_type)
