"""Cycle-level decomposition of the CTA-pair linear kernel (CMDI_TEST_DBG counters)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402
from tools.gpu_probe import probe_linear  # noqa: E402

os.environ["CMDI_TEST_DBG"] = "1"
os.environ["CMDI_TEST_NOF32"] = "1"
lib = C.capi.load()
for dbg in ("0", "8", "1", "2", "10", "3"):
    os.environ["CMDI_DEBUG"] = dbg
    print("CMDI_DEBUG =", dbg, "(1 skip store blocks, 2 skip mma, 8 staging without the STG)", flush=True)
    probe_linear(lib, 12608, 1536, 512, 3, -256)
