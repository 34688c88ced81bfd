"""A few B=64 sampling steps of the MDM_UNET (xl geometry, keyframe-conditioned, no CFG) for ncu captures / timing."""
import os, sys, time
os.environ.setdefault("CMDI_NO_GRAPH", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
dev = torch.device("cuda:0")
B = 64
m = C.MDM_UNET(keyframe_conditioned=True, zero=False).to(dev)
xo = torch.randn(B, 263, 1, 196, device=dev)
kf = C.get_keyframes_mask(xo, torch.full((B,), 196), "benchmark_sparse", trans_length=5)
d = C.create_gaussian_diffusion()
d.rng = "engine"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kw = {"y": {}, "obs_x0": xo, "obs_mask": kf}
d.p_sample_loop(m, (B, 263, 1, 196), model_kwargs=kw, skip_timesteps=1000 - n)
torch.cuda.synchronize()
t0 = time.perf_counter()
d.p_sample_loop(m, (B, 263, 1, 196), model_kwargs=kw, skip_timesteps=1000 - n)
torch.cuda.synchronize()
print(f"unet xl B=64 no-CFG: {n / (time.perf_counter() - t0):.1f} steps/s")
