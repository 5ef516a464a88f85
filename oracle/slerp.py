"""ORACLE (test infrastructure only).  Restatement of `slerp` (reference stable_diffusion_videos/utils.py:42-66):
numpy, computed in the dtype of the inputs, dot over the whole flattened array, lerp fallback above 0.9995.
`slerp_f64` is the fp64 ground truth the KATs compare both against."""
import numpy as np
import torch


def slerp(t, v0, v1, DOT_THRESHOLD=0.9995):
    is_torch = isinstance(v0, torch.Tensor)
    if is_torch:
        dev = v0.device
        v0 = v0.cpu().numpy()
        v1 = v1.cpu().numpy()
    dot = np.sum(v0 * v1 / (np.linalg.norm(v0) * np.linalg.norm(v1)))  # utils.py:50
    if np.abs(dot) > DOT_THRESHOLD:
        v2 = (1 - t) * v0 + t * v1  # utils.py:52
    else:
        theta_0 = np.arccos(dot)
        sin_theta_0 = np.sin(theta_0)
        theta_t = theta_0 * t
        sin_theta_t = np.sin(theta_t)
        s0 = np.sin(theta_0 - theta_t) / sin_theta_0
        s1 = sin_theta_t / sin_theta_0
        v2 = s0 * v0 + s1 * v1  # utils.py:60
    if is_torch:
        v2 = torch.from_numpy(np.asarray(v2)).to(dev)
    return v2


def slerp_f64(t, v0, v1, DOT_THRESHOLD=0.9995):
    a = np.asarray(v0, dtype=np.float64)
    b = np.asarray(v1, dtype=np.float64)
    return slerp(float(t), a, b, DOT_THRESHOLD)
