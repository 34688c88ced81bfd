"""Smallest end-to-end check: one denoiser pass (B=2) against the CPU oracle. Used under compute-sanitizer."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402
from oracle import condmdi_oracle as O  # noqa: E402

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sd = O.random_state_dict(seed=7, layers=layers)
m = C.MDM(num_layers=layers)
m.load_state_dict(sd, strict=False)
m = m.cuda()
x = O.golden_inputs()["x"]
t = torch.tensor([37, 37])
got = m(x.cuda(), t.cuda(), y={})
torch.cuda.synchronize()
ref = O.mdm_forward(sd, x, t)
err = (got.cpu().double() - ref.double()).abs()
print(f"forward layers={layers}: max_abs={err.max():.3e} mean_abs={err.mean():.3e} ref_absmax={ref.abs().max():.3e}")
