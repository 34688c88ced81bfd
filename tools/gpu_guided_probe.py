"""A few reconstruction-guided steps (config 4, CFG, B=64) for an ncu launch list of the backward pass:
ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c 260 --csv python tools/gpu_guided_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402

B, D, L = 64, 263, 196
dev = torch.device("cuda:0")
torch.manual_seed(0)
text = C.MDM(cond_mode="text", cond_mask_prob=0.1).to(dev)
cond = torch.randn(B, 512, device=dev)
text.encode_text = lambda t: cond
cfg = C.ClassifierFreeSampleModel(text)
x_obs = torch.randn(B, D, 1, L, device=dev)
kf = C.get_keyframes_mask(x_obs, torch.full((B,), 196), "benchmark_sparse", trans_length=5)
y = {"text": [""] * B, "text_scale": torch.full((B,), 2.5, device=dev), "mask": torch.ones(B, 1, 1, L, dtype=torch.bool, device=dev),
     "imputate": 1, "stop_imputation_at": 1, "replacement_distribution": "conditional", "inpainted_motion": x_obs,
     "inpainting_mask": kf, "reconstruction_guidance": True, "reconstruction_weight": 20.0, "gradient_schedule": None,
     "diffusion_steps": 1000, "stop_recguidance_at": 0}
diff = C.create_gaussian_diffusion()
diff.use_graph = False   # plain launches: every kernel shows up by name in the launch list
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
out = diff.p_sample_loop(cfg, (B, D, 1, L), model_kwargs={"y": y}, skip_timesteps=1000 - n)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
