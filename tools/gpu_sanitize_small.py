"""Small invocations of the newer kernels for compute-sanitizer (memcheck): persistent attention, tensor-core attention
backward, ATen-compatible normal stream, recover_from_ric, and a 3-step guided loop.
    compute-sanitizer --tool memcheck python tools/gpu_sanitize_small.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402

lib = C.capi.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
g = torch.Generator(device="cuda").manual_seed(0)
for nseq, S, H in ((3, 197, 4), (2, 129, 2), (2, 5, 2)):
    qkv = torch.randn(nseq * S, 3 * H * 128, device="cuda", generator=g)
    out = torch.empty(nseq * S, H * 128, device="cuda")
    C.capi.check(lib.cmdi_test_attention(p(qkv), p(out), nseq, S, H, 3, None))
    dO = torch.randn(nseq * S, H * 128, device="cuda", generator=g)
    dq = torch.empty_like(qkv)
    C.capi.check(lib.cmdi_test_attention_bwd(p(qkv), p(dO), p(dq), nseq, S, H, None))
torch.cuda.synchronize()
n = 4 * 263 * 196
z = torch.empty(n, device="cuda")
C.capi.check(lib.cmdi_test_normal_aten(p(z), n, 1234, 8, 256 * 806, None))
x = torch.randn(3, 263, 1, 196, device="cuda")
j = C.sample_to_joints(x, torch.zeros(263), torch.ones(263), 22, False)
j2 = C.recover_from_ric(torch.randn(2, 1, 57, 263, device="cuda"), 22, True)
torch.cuda.synchronize()
# a short guided loop through the public API (forward with stash + tensor-core backward + guided step kernel)
B, D, L = 2, 263, 196
m = C.MDM(cond_mode="text", cond_mask_prob=0.1).cuda()
cond = torch.randn(B, 512, device="cuda")
m.encode_text = lambda t: cond
cfg = C.ClassifierFreeSampleModel(m)
x_obs = torch.randn(B, D, 1, L, device="cuda")
kf = C.get_keyframes_mask(x_obs, torch.full((B,), 196), "benchmark_sparse", trans_length=5)
y = {"text": [""] * B, "text_scale": torch.full((B,), 2.5, device="cuda"), "mask": torch.ones(B, 1, 1, L, dtype=torch.bool, device="cuda"),
     "imputate": 1, "stop_imputation_at": 1, "replacement_distribution": "conditional", "inpainted_motion": x_obs,
     "inpainting_mask": kf, "reconstruction_guidance": True, "reconstruction_weight": 20.0, "gradient_schedule": None,
     "diffusion_steps": 1000, "stop_recguidance_at": 0}
diff = C.create_gaussian_diffusion()
diff.use_graph = False
out = diff.p_sample_loop(cfg, (B, D, 1, L), model_kwargs={"y": y}, skip_timesteps=997)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()), float(j.abs().mean()), float(j2.abs().mean()), float(z.std()))
# round 2: the chained transformer path (LayerNorm folded, dependency counters) and a small MDM_UNET, plain launches
plain = C.MDM(num_layers=2).cuda()
d2 = C.create_gaussian_diffusion(timestep_respacing="ddim50")
d2.use_graph = False
d2.rng = "engine"
o1 = d2.ddim_sample_loop(plain, (3, D, 1, L), model_kwargs={"y": {}}, skip_timesteps=47)
unet = C.MDM_UNET(dim_mults=(1, 1), keyframe_conditioned=True, zero=False).cuda()
o2 = d2.ddim_sample_loop(unet, (B, D, 1, L), model_kwargs={"y": {}, "obs_x0": x_obs, "obs_mask": kf}, skip_timesteps=48)
torch.cuda.synchronize()
print("ok round 2", float(o1.abs().mean()), float(o2.abs().mean()))
