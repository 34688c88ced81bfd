#!/usr/bin/env python
"""PyTorch-eager baseline of the same denoising step on the GPU ("the library kernels to beat").

Builds the MDM trans_enc architecture from stock torch.nn modules (nn.TransformerEncoder, post-norm, erf-GELU,
8L/512d/4h/ff1024 -- what model/mdm.py:107-114 instantiates), random-init, fp32, and times one DDPM loop iteration
(one forward + the posterior/noise update written with torch ops) at B=64, L=196, D=263 with CUDA events.
Reported for fp32 (torch default: TF32 off, the reference's numerics) and with TF32 matmuls allowed.

    python tools/gpu_eager_baseline.py [steps]
"""
import json
import math
import sys

import torch
import torch.nn as nn

B, D, L, DM, FF, LAYERS, HEADS, T = 64, 263, 196, 512, 1024, 8, 4, 1000


class EagerMDM(nn.Module):
    def __init__(self):
        super().__init__()
        self.pose = nn.Linear(D, DM)
        layer = nn.TransformerEncoderLayer(d_model=DM, nhead=HEADS, dim_feedforward=FF, dropout=0.1, activation="gelu")
        self.enc = nn.TransformerEncoder(layer, num_layers=LAYERS)
        self.time = nn.Sequential(nn.Linear(DM, DM), nn.SiLU(), nn.Linear(DM, DM))
        self.final = nn.Linear(DM, D)
        pe = torch.zeros(5000, DM)
        pos = torch.arange(0, 5000, dtype=torch.float).unsqueeze(1)
        div = torch.exp(torch.arange(0, DM, 2).float() * (-math.log(10000.0) / DM))
        pe[:, 0::2], pe[:, 1::2] = torch.sin(pos * div), torch.cos(pos * div)
        self.register_buffer("pe", pe.unsqueeze(1))

    def forward(self, x, t):
        emb = self.time(self.pe[t])                                   # (B,1,d)
        h = self.pose(x.permute(3, 0, 1, 2).reshape(L, -1, D))        # (L,B,d)
        seq = torch.cat((emb.permute(1, 0, 2), h), 0)
        seq = seq + self.pe[: seq.shape[0]]
        out = self.enc(seq)[1:]
        return self.final(out).reshape(L, -1, D, 1).permute(1, 2, 3, 0)


def step(model, x, t, coef1, coef2, logvar):
    x0 = model(x, t)
    mean = coef1[t].view(-1, 1, 1, 1) * x0 + coef2[t].view(-1, 1, 1, 1) * x
    noise = torch.randn_like(x)
    nz = (t != 0).float().view(-1, 1, 1, 1)
    return mean + nz * torch.exp(0.5 * logvar[t].view(-1, 1, 1, 1)) * noise


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = EagerMDM().to(dev).eval()
    coef1, coef2, logvar = (torch.rand(T, device=dev) for _ in range(3))
    out = {}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        x = torch.randn(B, D, 1, L, device=dev)
        with torch.no_grad():
            for i in range(5):
                x = step(model, x, torch.full((B,), T - 1 - i, device=dev), coef1, coef2, logvar)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                x = step(model, x, torch.full((B,), T - 6 - i, device=dev), coef1, coef2, logvar)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out[f"torch_eager_{name}_steps_per_s"] = round(1e3 / ms, 2)
        out[f"torch_eager_{name}_ms_per_step"] = round(ms, 3)
    out["note"] = "stock torch.nn.TransformerEncoder (eval fast path), random-init, B=64 L=196 D=263, one DDPM step per iteration"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
