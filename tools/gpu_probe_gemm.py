"""GEMM bring-up probe: CTA-pair kernel correctness + a time decomposition (loads / MMAs / stores) of every linear."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import condmdi_b200 as C  # noqa: E402
from tools.gpu_probe import probe_linear  # noqa: E402


def decomposition():
    torch.manual_seed(0)
    model = C.MDM().cuda()
    sd = {k: v for k, v in model.state_dict().items()}
    rows = []
    for gemm in ("single", "pair"):
        for dbg, label in ((0, "full"), (1, "no-store"), (2, "no-mma"), (4, "no-load"), (5, "mma-only"), (3, "load-only")):
            os.environ["CMDI_GEMM"], os.environ["CMDI_DEBUG"] = gemm, str(dbg)
            eng = C.Engine(torch.device("cuda:0"), max_batch=64, precision=int(os.environ.get("PREC", "3")))
            eng.load_state_dict(sd)
            eng.profile_pass(64)
            agg = {}
            for _ in range(3):
                for name, ms in eng.profile_pass(64):
                    agg[name] = agg.get(name, 0.0) + ms / 3
            eng.close()
            rows.append((gemm, label, agg))
            print(f"{gemm:6s} {label:9s} " + " ".join(f"{k}={agg[k] / (8 if k not in ('frame_embed', 'out_head', 'token_rows') else 1) * 1e3:7.1f}us" for k in
                                                       ("qkv", "out_proj", "ffn1", "ffn2", "attention", "ln1", "frame_embed", "out_head")), flush=True)
    os.environ.pop("CMDI_DEBUG")
    os.environ.pop("CMDI_GEMM")


def main():
    lib = C.capi.load()
    print(torch.cuda.get_device_name(0), flush=True)
    if len(sys.argv) > 1 and sys.argv[1] == "decomp":
        decomposition()
        return
    for (M, N, K, prec, bn, kw) in [
        (256, 256, 64, 1, -256, dict(use_bias=False)), (256, 256, 64, 3, -256, dict(use_bias=False)), (256, 128, 64, 3, -128, dict(use_bias=False)),
        (128, 256, 512, 3, -256, {}), (1000, 1536, 512, 3, -256, {}), (12608, 1536, 512, 3, -256, {}),
        (12608, 1024, 512, 3, -256, dict(act=1)), (12608, 512, 1024, 3, -128, dict(use_res=True)), (12544, 512, 263, 3, -128, {}),
        (12608, 264, 512, 3, -128, {}), (25216, 1536, 512, 3, -256, {}),
    ]:
        probe_linear(lib, M, N, K, prec, bn, **kw)


if __name__ == "__main__":
    main()
