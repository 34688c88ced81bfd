"""Throughput of the BASELINE.json configs 2-5 at B=64 on one B200 (synthetic inputs, random-init weights)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402

B, D, L = 64, 263, 196
dev = torch.device("cuda:0")
torch.manual_seed(0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100


def timed(fn, n):
    fn(3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn(n)
    e1.record()
    torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) * 1e-3)


plain = C.MDM(cond_mode="no_cond").to(dev)
text = C.MDM(cond_mode="text", cond_mask_prob=0.1).to(dev)
cond = torch.randn(B, 512, generator=torch.Generator().manual_seed(1)).to(dev)
text.encode_text = lambda t: cond
cfg = C.ClassifierFreeSampleModel(text)
x_obs = torch.randn(B, D, 1, L, device=dev)
lengths = torch.full((B,), 196)
kf = C.get_keyframes_mask(x_obs, lengths, "benchmark_sparse", trans_length=5)
y_mask = torch.ones(B, 1, 1, L, dtype=torch.bool, device=dev)
scale = torch.full((B,), 2.5, device=dev)
d1000 = C.create_gaussian_diffusion()
d100 = C.create_gaussian_diffusion(use_ddim=True)
out = {}


def run(diff, model, y, sampler="p_sample_loop", T=1000):
    def fn(n):
        getattr(diff, sampler)(model, (B, D, 1, L), model_kwargs={"y": y}, skip_timesteps=T - n)
    return fn


out["config2_uncond_ddpm"] = timed(run(d1000, plain, {}), steps)
y3 = {"text": [""] * B, "text_scale": scale, "mask": y_mask, "imputate": 1, "stop_imputation_at": 1,
      "replacement_distribution": "conditional", "inpainted_motion": x_obs, "inpainting_mask": kf}
out["config3_cfg2.5_sparse_keyframe_imputation"] = timed(run(d1000, cfg, y3), steps)
y4 = dict(y3, reconstruction_guidance=True, reconstruction_weight=20.0, gradient_schedule=None, diffusion_steps=1000, stop_recguidance_at=0)
out["config4_cfg_imputation_recon_guidance_w20"] = timed(run(d1000, cfg, y4), max(10, steps // 5))
y4b = {k: v for k, v in y4.items() if k not in ("text", "text_scale")}
out["config4b_no_cfg_imputation_recon_guidance_w20"] = timed(run(d1000, plain, y4b), max(10, steps // 5))
out["config5_ddim100_per_gpu"] = timed(run(d100, plain, {}, "ddim_sample_loop", T=100), min(steps, 100))
print(json.dumps({k: round(v, 2) for k, v in out.items()}))
