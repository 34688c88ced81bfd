"""HumanML3D feature vectors -> joint positions on the GPU.

Mirrors `data_loaders/humanml/scripts/motion_process.py:474-489` (`recover_from_ric`, with
`recover_root_rot_pos` :402-441) of the reference and the CPU hop around it in `sample/synthesize.py:153-157`
(`sample.cpu().permute(0, 2, 3, 1)` -> `inv_transform` -> `recover_from_ric` -> `view/permute`).  The work runs in
`cmdi_recover_from_ric` (csrc/elementwise.cu); there is no CPU path.
"""
from __future__ import annotations

import torch

from . import capi
from .engine import _ptr, _stream_ptr


def _check(data: torch.Tensor):
    if not data.is_cuda:
        raise RuntimeError("condmdi_b200.recover_from_ric runs on CUDA tensors only (no CPU fallback)")


def recover_from_ric(data: torch.Tensor, joints_num: int, abs_3d: bool = False) -> torch.Tensor:
    """Same contract as the reference function: `data` (..., nframes, nfeats) de-normalised HumanML3D vectors ->
    (..., nframes, joints_num, 3) positions (root first)."""
    _check(data)
    lead, (L, C) = data.shape[:-2], data.shape[-2:]
    x = data.to(torch.float32).reshape(-1, L, C).contiguous()
    out = torch.empty(x.shape[0], L, joints_num, 3, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(capi.load().cmdi_recover_from_ric(_ptr(x), L * C, C, 1, None, None, x.shape[0], L, C, joints_num, int(abs_3d),
                                                     _ptr(out), L * joints_num * 3, joints_num * 3, 3, 1, _stream_ptr(x.device)),
                   "cmdi_recover_from_ric")
    return out.reshape(*lead, L, joints_num, 3)


def sample_to_joints(sample: torch.Tensor, mean, std, joints_num: int = 22, abs_3d: bool = False) -> torch.Tensor:
    """What sample/synthesize.py:153-157 computes from the sampler output, without leaving the GPU:
    `sample` (B, nfeats, 1, nframes) normalised -> (B, joints_num, 3, nframes) positions, with
    `mean`/`std` (nfeats,) the dataset statistics of `t2m_dataset.inv_transform`."""
    _check(sample)
    B, C, one, L = sample.shape
    assert one == 1
    x = sample.to(torch.float32).contiguous()
    mean_t = torch.as_tensor(mean, dtype=torch.float32).to(x.device).contiguous()
    std_t = torch.as_tensor(std, dtype=torch.float32).to(x.device).contiguous()
    assert mean_t.shape == std_t.shape == (C,)
    out = torch.empty(B, joints_num, 3, L, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        capi.check(capi.load().cmdi_recover_from_ric(_ptr(x), C * L, 1, L, _ptr(mean_t), _ptr(std_t), B, L, C, joints_num, int(abs_3d),
                                                     _ptr(out), joints_num * 3 * L, 1, 3 * L, L, _stream_ptr(x.device)),
                   "cmdi_recover_from_ric")
    return out
