"""Keyframe masks and guidance schedule of the reference's utils/editing_util.py, vectorised (no Python loops
over the batch) and device-agnostic (runs wherever `data` lives; integer/bool arithmetic only, bit-exact).

    get_keyframes_mask       utils/editing_util.py:56-229   (inference modes: benchmark_sparse :85-91,
                                                              benchmark_clip :93-100, uncond :102-105)
    joint_to_full_mask       utils/editing_util.py:30-44 with the 22x263 incidence matrices of
                             data_loaders/humanml_utils.py:68-91
    get_gradient_schedule    utils/editing_util.py:299-322
"""
from __future__ import annotations

import numpy as np
import torch


def _incidence(device) -> dict:
    pos = torch.zeros(22, 263, dtype=torch.bool)
    pos[0, 1:4] = True
    rot = torch.zeros(22, 263, dtype=torch.bool)
    rot[0, 0] = True
    vel = torch.zeros(22, 263, dtype=torch.bool)
    cnt = torch.zeros(22, 263, dtype=torch.bool)
    for j in range(1, 22):
        pos[j, 4 + 3 * j - 3:4 + 3 * j] = True
        rot[j, 4 + 63 + 6 * j - 6:4 + 63 + 6 * j] = True
    for j in range(22):
        vel[j, 4 + 63 + 126 + 3 * j:4 + 63 + 126 + 3 * (j + 1)] = True
    cnt[7, -4] = cnt[10, -3] = cnt[8, -2] = cnt[11, -1] = True
    return {k: v.to(device) for k, v in dict(pos=pos, rot=rot, vel=vel, cnt=cnt).items()}


def joint_to_full_mask(joint_mask: torch.Tensor, mode: str = "pos_rot_vel") -> torch.Tensor:
    """(B, 22, 1, L) bool -> (B, 263, 1, L) bool: a feature is observed iff one of its joints is."""
    assert mode in ["pos", "pos_rot", "pos_rot_vel"]
    inc = _incidence(joint_mask.device)
    m = inc["pos"] | inc["cnt"]
    if mode in ["pos_rot", "pos_rot_vel"]:
        m = m | inc["rot"]
    if mode == "pos_rot_vel":
        m = m | inc["vel"]
    # out[b, f, 0, l] = any_j joint_mask[b, j, 0, l] & m[j, f]: counts <= 22, exact in fp32 (the reference's
    # bool_matmul does the same float matmul and asserts exactness, editing_util.py:8-11)
    jm = joint_mask[:, :, 0, :].to(torch.float32)                     # (B, 22, L)
    out = torch.einsum("bjl,jf->bfl", jm, m.to(torch.float32)) > 0.5  # (B, 263, L)
    return out.unsqueeze(2)


def get_keyframes_mask(data, lengths, edit_mode="benchmark_sparse", trans_length=10, feature_mode="pos_rot_vel",
                       get_joint_mask=False, n_keyframes=5):
    batch_size, n_joints, n_features, n_frames = data.shape
    if n_joints != 263:
        raise ValueError("Unknown number of joints: {}".format(n_joints))  # the AMASS (764) branch is out of scope
    dev = data.device
    lengths = torch.as_tensor(lengths, device=dev).to(torch.int64).reshape(-1, 1)
    f = torch.arange(n_frames, device=dev).reshape(1, -1)
    if edit_mode == "benchmark_sparse":
        frame_obs = (f < lengths) & (f % trans_length == 0)                     # range(length)[::trans_length]
    elif edit_mode == "benchmark_clip":
        end = torch.div(lengths - trans_length, 2, rounding_mode="floor")       # (length - trans_length) // 2
        frame_obs = (f < end) | ((f >= end + trans_length) & (f < lengths))
    elif edit_mode == "uncond":
        frame_obs = torch.zeros(batch_size, n_frames, dtype=torch.bool, device=dev)
    else:
        raise NotImplementedError(f"edit_mode {edit_mode!r}: only the inference-time benchmark modes are provided")
    obs_joint_mask = frame_obs[:, None, None, :].expand(batch_size, 22, n_features, n_frames).contiguous()
    obs_feature_mask = joint_to_full_mask(obs_joint_mask, mode=feature_mode)
    if get_joint_mask:
        return obs_feature_mask, obs_joint_mask
    return obs_feature_mask


def get_gradient_schedule(schedule_name=None, num_diffusion_steps=1000, scale=.05):
    if schedule_name is None:
        return np.ones(num_diffusion_steps)
    if schedule_name == 'first-half':
        return np.concatenate((np.ones(num_diffusion_steps // 2), np.zeros(num_diffusion_steps - num_diffusion_steps // 2)))
    if schedule_name == 'last-half':
        return np.concatenate((np.zeros(num_diffusion_steps // 2), np.ones(num_diffusion_steps // 2)))
    if schedule_name == 'exponential':
        ts = np.arange(num_diffusion_steps)[::-1]
        return np.exp(-scale * ts)
    elif schedule_name == 'sigmoid':
        ts = np.arange(num_diffusion_steps)
        scale /= 5
        return 1 / (1 + np.exp(scale * (-ts + num_diffusion_steps / 2)))
    elif schedule_name == 'half-sigmoid':
        ts = np.arange(num_diffusion_steps)
        scale /= 5
        return 1 / (1 + np.exp(scale * (-ts)))
    raise NotImplementedError(f"unknown guidance schedule for reconstruction guidance: {schedule_name}")


def requires_imputation(model_kwargs, denoising_step):
    """utils/editing_util.py:336-346 (host predicate; the engine evaluates the same test on the device)."""
    y = model_kwargs['y']
    if 'imputate' not in y.keys():
        return False
    if y['imputate']:
        assert 'stop_imputation_at' in y.keys()
        assert 'inpainting_mask' in y.keys() and 'inpainted_motion' in y.keys()
        return bool((denoising_step >= y['stop_imputation_at']).all())
    return False


def requires_reconstruction_guidance(model_kwargs, denoising_step):
    """utils/editing_util.py:325-333."""
    y = model_kwargs['y']
    if 'reconstruction_guidance' not in y.keys():
        return False
    if y['reconstruction_guidance']:
        assert 'stop_recguidance_at' in y.keys()
        assert 'inpainting_mask' in y.keys() and 'inpainted_motion' in y.keys()
        return bool((denoising_step >= y['stop_recguidance_at']).all())
    return False
