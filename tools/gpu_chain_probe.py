"""First light of the chained forward path: parity of one forward pass (B=2 and B=64) vs the CPU oracle, then timing."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
from oracle import condmdi_oracle as O

dev = torch.device("cuda:0")
sd = O.random_state_dict(seed=7)
m = C.MDM()
m.load_state_dict(sd, strict=False)
m = m.to(dev)
g = torch.Generator().manual_seed(3)
for B in (2, 64):
    x = torch.randn(B, 263, 1, 196, generator=g)
    t = torch.full((B,), 123)
    got = m(x.to(dev), t.to(dev), y={}).cpu()
    want = O.mdm_forward(sd, x[:4], t[:4])
    err = (got[:4] - want).abs()
    viol = (err > 1e-4 + 1e-3 * want.abs()).float().mean().item()
    print(f"B={B}: max_abs={err.max():.3e} mean_abs={err.mean():.3e} viol={viol:.2e} finite={bool(torch.isfinite(got).all())}", flush=True)
eng = m.engine_for(dev, max_batch=64)
d = C.create_gaussian_diffusion()
eng.set_schedule(d.betas, d.timestep_map)
for name, ms in eng.profile_pass(64)[:7]:
    print(f"  {name}: {ms*1e3:.1f} us")
eng.sample(64, num_steps=20, seed=1)
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.sample(64, num_steps=200, seed=1)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"B=64: {200/dt:.1f} steps/s ({dt/200*1e3:.3f} ms/step)")
