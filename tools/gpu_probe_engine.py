"""Engine bring-up probe (GPU box): denoiser pass and short loops against the CPU oracle, plus a first timing."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import condmdi_b200 as C  # noqa: E402
from oracle import condmdi_oracle as O  # noqa: E402


def stats(name, got, ref):
    got, ref = got.double().cpu(), ref.double().cpu()
    err = (got - ref).abs()
    viol = (err > (1e-4 + 1e-3 * ref.abs())).double().mean().item()
    print(f"{name:55s} max_abs={err.max().item():.3e} mean_abs={err.mean().item():.3e} ref_absmax={ref.abs().max().item():.3e} "
          f"viol={viol:.5f} nan={int(torch.isnan(got).sum())}", flush=True)


def make_model(sd, text):
    m = C.MDM(cond_mode="text" if text else "no_cond", cond_mask_prob=0.1)
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    return m.cuda()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    dev = torch.device("cuda:0")
    gi = O.golden_inputs()
    x, cond, x_obs, tape, scale, lengths, y_mask, kf = (gi[k] for k in ("x", "cond", "x_obs", "tape", "text_scale", "lengths", "y_mask", "kf_mask"))
    B, D, L = 2, 263, 196
    sd = O.random_state_dict(seed=7, text=False)
    sdt = O.random_state_dict(seed=7, text=True)
    m = make_model(sd, False)
    mt = make_model(sdt, True)
    mt.encode_text = lambda texts: cond.to(dev)

    # ---- single pass ----
    for tv in (999, 37):
        t = torch.tensor([tv, tv])
        ref = O.mdm_forward(sd, x, t)
        got = m(x.to(dev), t.to(dev), y={})
        stats(f"MDM.forward no_cond t={tv}", got, ref)
    t = torch.tensor([500, 500])
    stats("MDM.forward text", mt(x.to(dev), t.to(dev), y={"text": ["a", "b"]}), O.mdm_forward(sdt, x, t, cond))
    stats("MDM.forward uncond", mt(x.to(dev), t.to(dev), y={"text": ["a", "b"], "uncond": True}), O.mdm_forward(sdt, x, t, cond, uncond=True))
    cfgm = C.ClassifierFreeSampleModel(mt)
    stats("CFG forward", cfgm(x.to(dev), t.to(dev), y={"text": ["a", "b"], "text_scale": scale.to(dev)}), O.cfg_forward(sdt, x, t, cond, scale))

    # ---- loops with a shared tape ----
    diff = C.create_gaussian_diffusion()
    diff.noise_tape = tape.to(dev)
    tab = O.make_tables("")
    eng = m.engine_for(dev, max_batch=2)
    eng.set_schedule(diff.betas, diff.timestep_map)
    for use_graph in (False, True):
        res = eng.sample(2, x_T=tape[0].to(dev), noise_tape=tape[1:].to(dev), num_steps=3, want_pred_xstart=True, use_graph=use_graph)
        ref = O.sample_loop(sd, tab, (B, D, 1, L), O.Conditioning(), tape, "ddpm", max_steps=3, return_all=True)
        stats(f"ddpm uncond 3 steps sample (graph={use_graph})", res["sample"], ref[-1]["sample"])
        stats(f"ddpm uncond 3 steps pred_xstart (graph={use_graph})", res["pred_xstart"], ref[-1]["pred_xstart"])

    d50 = C.create_gaussian_diffusion(timestep_respacing="ddim50")
    tape50 = tape[torch.arange(51) % 8]
    d50.noise_tape = tape50.to(dev)
    got = d50.ddim_sample_loop(m, (B, D, 1, L), model_kwargs={"y": {}})
    ref = O.sample_loop(sd, O.make_tables("ddim50"), (B, D, 1, L), O.Conditioning(), tape50, "ddim")
    stats("ddim50 full loop", got, ref)

    ykw = {"text": ["a", "b"], "text_scale": scale.to(dev), "mask": y_mask.to(dev), "lengths": lengths, "imputate": 1,
           "stop_imputation_at": 1, "replacement_distribution": "conditional", "inpainted_motion": x_obs.to(dev),
           "inpainting_mask": kf.to(dev)}
    got = diff.p_sample_loop(cfgm, (B, D, 1, L), model_kwargs={"y": ykw}, skip_timesteps=996, init_image=x_obs.to(dev))
    c = O.Conditioning(cond_emb=cond, cfg=True, text_scale=scale, y_mask=y_mask, imputate=True, stop_imputation_at=1,
                       inpainted_motion=x_obs, inpainting_mask=kf)
    ref = O.sample_loop(sdt, tab, (B, D, 1, L), c, tape, "ddpm", skip_timesteps=996, init_image=x_obs)
    stats("cfg + imputation, last 4 steps", got, ref)

    # ---- timing at the benchmark shape ----
    for prec, name in ((C.capi.PRECISION_BF16X3, "bf16x3"), (C.capi.PRECISION_BF16, "bf16")):
        e = m.engine_for(dev, max_batch=64, precision=prec)
        e.set_schedule(diff.betas, diff.timestep_map)
        for steps in (5, 50):
            e.sample(64, seed=1, num_steps=steps)
            torch.cuda.synchronize()
            t0 = time.time()
            e.sample(64, seed=1, num_steps=steps)
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(f"B=64 {name} {steps} steps: {dt * 1e3:.2f} ms -> {steps / dt:.1f} steps/s", flush=True)
    print("launches", e.launch_count)


if __name__ == "__main__":
    main()
