import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
from tools.gpu_probe import probe_attention
os.environ["CMDI_TEST_DBG"] = "1"
lib = C.capi.load()
probe_attention(lib, 64, 197, 4, 3)
probe_attention(lib, 64, 197, 4, 1)
probe_attention(lib, 8, 197, 4, 3)
