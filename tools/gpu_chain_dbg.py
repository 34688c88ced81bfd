"""Cycle counters of the chained linear kernel (CMDI_CHAIN_DBG=1), B=64."""
import os, sys
os.environ.setdefault("CMDI_CHAIN_DBG", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
dev = torch.device("cuda:0")
m = C.MDM().to(dev)
eng = m.engine_for(dev, max_batch=64)
d = C.create_gaussian_diffusion()
eng.set_schedule(d.betas, d.timestep_map)
eng.profile_pass(64, repeats=1)
prof = eng.profile_pass(64, repeats=1)
print({k: round(v * 1e3, 1) for k, v in prof[:7]})
