"""A few B=64 sampling steps (plain launches, no graph) for ncu captures: python tools/gpu_b64_steps.py [steps]"""
import os, sys
os.environ.setdefault("CMDI_NO_GRAPH", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
dev = torch.device("cuda:0")
m = C.MDM().to(dev)
eng = m.engine_for(dev, max_batch=64)
d = C.create_gaussian_diffusion()
eng.set_schedule(d.betas, d.timestep_map)
eng.sample(64, num_steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3, seed=1)
torch.cuda.synchronize()
