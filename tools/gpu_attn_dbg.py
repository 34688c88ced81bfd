"""Phase counters of the attention kernel (CMDI_TEST_DBG=1) at the BASELINE shape: 64 sequences x 197 tokens x 4 heads."""
import ctypes, os, sys
os.environ.setdefault("CMDI_TEST_DBG", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C
lib = C.capi.load()
p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
nseq, S, H = 64, 197, 4
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn(nseq * S, 3 * H * 128, device="cuda", generator=g)
out = torch.empty(nseq * S, H * 128, device="cuda")
for _ in range(2):
    C.capi.check(lib.cmdi_test_attention(p(qkv), p(out), nseq, S, H, 3, None))
torch.cuda.synchronize()
