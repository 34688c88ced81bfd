"""First light of the MDM_UNET path: forward parity vs tests/golden/unet.npz (reference outputs) and a small config vs the oracle."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
from oracle import condmdi_oracle as O

dev = torch.device("cuda:0")
gi = O.golden_inputs()
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "unet.npz"))

def rep(a, b, what):
    a, b = torch.as_tensor(a).cpu().double(), torch.as_tensor(b).cpu().double()
    e = (a - b).abs()
    viol = (e > 1e-4 + 1e-3 * b.abs()).double().mean().item()
    print(f"[{what}] max_abs={e.max():.3e} mean_abs={e.mean():.3e} ref_absmax={b.abs().max():.2f} viol={viol:.2e}", flush=True)

# small configuration first (fast to debug): dim 512 x (1, 1), keyframe-conditioned, no text
for mults in ((1, 1), (1, 1, 1)):
    sd = O.random_unet_state_dict(seed=3, mults=mults)
    m = C.MDM_UNET(dim_mults=mults, keyframe_conditioned=True)
    m.load_state_dict(sd, strict=False)
    m = m.to(dev)
    t = torch.tensor([41, 41])
    got = m(gi["x"].to(dev), t.to(dev), y={}, obs_x0=gi["x_obs"].to(dev), obs_mask=gi["kf_mask"].to(dev))
    want = O.unet_forward(sd, gi["x"], t, None, False, gi["x_obs"], gi["kf_mask"])
    rep(got, want, f"unet {mults} forward vs oracle")

sd = O.random_unet_state_dict(seed=11, text=True)
m = C.MDM_UNET(keyframe_conditioned=True, cond_mode="text", cond_mask_prob=0.1)
m.load_state_dict(sd, strict=False)
m = m.to(dev)
_table = {"a": gi["cond"][0].to(dev), "b": gi["cond"][1].to(dev)}
m.encode_text = lambda texts: torch.stack([_table[t] for t in texts])  # per-sample timesteps are evaluated in groups
x, xo, kf = gi["x"].to(dev), gi["x_obs"].to(dev), gi["kf_mask"].to(dev)
got = m(x, torch.tensor(g["fwd.t"]).to(dev), y={"text": ["a", "b"]}, obs_x0=xo, obs_mask=kf)
rep(got, g["fwd.out"], "unet xl forward (text, keyframes) vs reference")
w = C.ClassifierFreeSampleModel(m)
got = w(x, torch.tensor([500, 500]).to(dev), y={"text": ["a", "b"], "text_scale": gi["text_scale"].to(dev)}, obs_x0=xo, obs_mask=kf)
rep(got, g["fwd_cfg.out"], "unet xl cfg forward vs reference")
d = C.create_gaussian_diffusion()
d.noise_tape = gi["tape"].to(dev)
kw = {"y": {"text": ["a", "b"], "text_scale": gi["text_scale"].to(dev), "mask": gi["y_mask"].to(dev), "lengths": gi["lengths"]},
      "obs_x0": xo, "obs_mask": kf}
got = d.p_sample_loop(w, (2, 263, 1, 196), model_kwargs=kw, skip_timesteps=996, init_image=xo)
rep(got, g["tail4.sample"], "unet xl ddpm 4-step tail vs reference")
# timing at B=64 (CFG: 128 sequences)
Bf = 64
cond = torch.randn(Bf, 512).to(dev)
m.encode_text = lambda texts: cond
xo64 = torch.randn(Bf, 263, 1, 196, device=dev)
kf64 = C.get_keyframes_mask(xo64, torch.full((Bf,), 196), "benchmark_sparse", trans_length=5)
kw = {"y": {"text": [""] * Bf, "text_scale": torch.full((Bf,), 2.5, device=dev), "mask": torch.ones(Bf, 1, 1, 196, dtype=torch.bool, device=dev)},
      "obs_x0": xo64, "obs_mask": kf64}
d = C.create_gaussian_diffusion()
d.rng = "engine"
d.p_sample_loop(w, (Bf, 263, 1, 196), model_kwargs=kw, skip_timesteps=995)
torch.cuda.synchronize()
t0 = time.perf_counter()
d.p_sample_loop(w, (Bf, 263, 1, 196), model_kwargs=kw, skip_timesteps=980)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"unet xl, B=64, CFG (2 passes/step): {20 / dt:.1f} steps/s ({dt / 20 * 1e3:.2f} ms/step)")
