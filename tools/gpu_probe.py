"""Kernel bring-up probe (run on the GPU box): each production kernel against an fp64 torch reference.

Prints one line per case with max-abs / relative error so a descriptor or layout mistake is visible
at a glance.  Not a test (tests/ holds those); this is the tool used while bringing kernels up.
"""
import ctypes
import os
import sys
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "diffusion-motion-inbetweening_b200", "libcondmdi_b200.so")


def load():
    lib = ctypes.CDLL(LIB)
    lib.cmdi_last_error.restype = ctypes.c_char_p
    return lib


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def stats(name, got, ref):
    got = got.double()
    err = (got - ref).abs()
    scale = ref.abs().max().item()
    viol = (err > (1e-4 + 1e-3 * ref.abs())).double().mean().item()
    print(f"{name:60s} max_abs={err.max().item():.3e} mean_abs={err.mean().item():.3e} ref_max={scale:.3e} "
          f"rel_max={err.max().item() / max(scale, 1e-30):.3e} viol(1e-3,1e-4)={viol:.4f} "
          f"nan={int(torch.isnan(got).sum())}", flush=True)


def probe_linear(lib, M, N, K, prec, block_n, act=0, use_bias=True, use_res=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g) if use_bias else None
    res = torch.randn(M, N, device="cuda", generator=g) if use_res else None
    C = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.cmdi_test_linear(ptr(A), ptr(W), ptr(bias), ptr(res), ptr(C), M, N, K, act, prec, block_n, None)
    torch.cuda.synchronize()
    if rc != 0:
        print(f"linear M={M} N={N} K={K} prec={prec} bn={block_n}: ERROR {lib.cmdi_last_error().decode()}")
        return
    ref = A.double() @ W.double().t()
    if bias is not None:
        ref = ref + bias.double()
    if res is not None:
        ref = ref + res.double()
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    stats(f"linear M={M} N={N} K={K} prec={prec} bn={block_n} act={act} res={int(use_res)}", C, ref)


def probe_attention(lib, nseq, S, H, prec, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(nseq * S, 3 * H * 128, device="cuda", generator=g)
    O = torch.full((nseq * S, H * 128), float("nan"), device="cuda")
    rc = lib.cmdi_test_attention(ptr(qkv), ptr(O), nseq, S, H, prec, None)
    torch.cuda.synchronize()
    if rc != 0:
        print(f"attention nseq={nseq} S={S} H={H} prec={prec}: ERROR {lib.cmdi_last_error().decode()}")
        return
    x = qkv.double().view(nseq, S, 3, H, 128)
    q, k, v = x[:, :, 0].transpose(1, 2), x[:, :, 1].transpose(1, 2), x[:, :, 2].transpose(1, 2)
    p = torch.softmax(q @ k.transpose(-1, -2) / 128 ** 0.5, dim=-1)
    ref = (p @ v).transpose(1, 2).reshape(nseq * S, H * 128)
    stats(f"attention nseq={nseq} S={S} H={H} prec={prec}", O, ref)


def probe_layernorm(lib, rows):
    g = torch.Generator(device="cuda").manual_seed(1)
    v = torch.randn(rows, 512, device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn(512, device="cuda", generator=g)
    beta = torch.randn(512, device="cuda", generator=g)
    out = torch.empty_like(v)
    rc = lib.cmdi_test_layernorm(ptr(v), ptr(gamma), ptr(beta), ptr(out), rows, None)
    torch.cuda.synchronize()
    if rc != 0:
        print("layernorm: ERROR", lib.cmdi_last_error().decode())
        return
    ref = torch.nn.functional.layer_norm(v.double(), (512,), gamma.double(), beta.double(), 1e-5)
    stats(f"layernorm rows={rows}", out, ref)


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda, flush=True)
    lib = load()
    cases = [
        lambda: probe_layernorm(lib, 1000),
        lambda: probe_linear(lib, 128, 128, 64, 1, 128, use_bias=False),
        lambda: probe_linear(lib, 128, 128, 64, 3, 128, use_bias=False),
        lambda: probe_linear(lib, 128, 256, 64, 1, 256, use_bias=False),
        lambda: probe_linear(lib, 256, 256, 512, 1, 128),
        lambda: probe_linear(lib, 256, 256, 512, 3, 128),
        lambda: probe_linear(lib, 256, 512, 512, 3, 256),
        lambda: probe_linear(lib, 1000, 1536, 512, 3, 256),
        lambda: probe_linear(lib, 12608, 1536, 512, 3, 256),
        lambda: probe_linear(lib, 12608, 1024, 512, 3, 256, act=1),
        lambda: probe_linear(lib, 12608, 512, 1024, 3, 128, use_res=True),
        lambda: probe_linear(lib, 12608, 512, 1024, 1, 128, use_res=True),
        lambda: probe_linear(lib, 12544, 512, 263, 3, 256),
        lambda: probe_linear(lib, 12608, 264, 512, 3, 128),
        lambda: probe_attention(lib, 1, 197, 1, 1),
        lambda: probe_attention(lib, 1, 197, 1, 3),
        lambda: probe_attention(lib, 3, 197, 4, 3),
        lambda: probe_attention(lib, 2, 100, 4, 3),
        lambda: probe_attention(lib, 64, 197, 4, 3),
    ]
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "case":
        sel = [cases[int(sys.argv[2])]]
    else:
        sel = {"ln": cases[:1], "linear": cases[1:14], "attention": cases[14:], "all": cases}[which]
    for c in sel:
        try:
            c()
        except Exception:  # keep going: a later case may still be informative
            traceback.print_exc()
            sys.stdout.flush()
            # a trapped kernel poisons the context: no point continuing in this process
            try:
                torch.cuda.synchronize()
            except Exception:
                print("CUDA context is dead; stopping", flush=True)
                break


if __name__ == "__main__":
    main()
