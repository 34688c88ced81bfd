"""A/B of the hi-plane split in the attention kernel (RN vs truncation): accuracy of the kernel alone."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
from tools.gpu_probe import probe_attention
lib = C.capi.load()
for mode in ("rn", "trunc"):
    os.environ["CMDI_ATTN_SPLIT"] = mode
    print("split =", mode)
    probe_attention(lib, 64, 197, 4, 3)
    probe_attention(lib, 3, 197, 4, 3, seed=5)
