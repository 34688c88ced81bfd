"""Drop-in mirror of the reference's sampler classes, backed by the native B200 engine.

Mirrors (same names, argument meaning and error behaviour; paths relative to the reference repository):
    get_named_beta_schedule / betas_for_alpha_bar   diffusion/gaussian_diffusion.py:24-71
    ModelMeanType / ModelVarType / DiffusionConfig  :74-136
    GaussianDiffusion  (sampling half)              :139-241, :311-349, :1149-1297, :1454-1587
    space_timesteps / SpacedDiffusion               diffusion/respace.py:9-62, :65-116
    create_gaussian_diffusion                       utils/model_util.py:122-165

`p_sample_loop` / `ddim_sample_loop` run the WHOLE loop in one native call (`cmdi_sample`): no per-step
Python, no per-step H2D table copies, no per-step host sync (the reference syncs on
`(t >= stop_imputation_at).all()`, utils/editing_util.py:344).  What the reference computes per step in
`p_mean_variance` / `p_sample` / `ddim_sample_with_grad` is done by the CUDA kernels in csrc/.

Not accelerated (raise NotImplementedError, like the reference does for its own unsupported branches):
cond_fn / 'gmd' classifier guidance, learned variances, EPSILON/PREVIOUS_X parametrisations,
const_noise.  `reconstruction_guidance` runs a backward pass through the denoiser on the GPU (csrc/backward.cu).
"""
from __future__ import annotations

import enum
import math
from dataclasses import dataclass, field
from typing import List, Optional

import warnings

import numpy as np
import torch

from . import capi
from .editing_util import get_gradient_schedule
from .model import resolve_model


def get_named_beta_schedule(schedule_name, num_diffusion_timesteps, scale_betas=1.):
    if schedule_name == "linear":
        scale = scale_betas * 1000 / num_diffusion_timesteps
        return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)
    elif schedule_name == "cosine":
        return betas_for_alpha_bar(num_diffusion_timesteps, lambda t: math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2)
    raise NotImplementedError(f"unknown beta schedule: {schedule_name}")


def betas_for_alpha_bar(num_diffusion_timesteps, alpha_bar, max_beta=0.999):
    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar(t2) / alpha_bar(t1), max_beta))
    return np.array(betas)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


@dataclass
class DiffusionConfig:
    betas: List = field(default_factory=list)
    model_mean_type: ModelMeanType = ModelMeanType.START_X
    model_var_type: ModelVarType = ModelVarType.FIXED_SMALL
    rescale_timesteps: bool = False
    # accepted for signature compatibility with the reference's DiffusionConfig (training-side options)
    extra: dict = field(default_factory=dict)


def _enum_name(v) -> str:
    return getattr(v, "name", str(v))


class GaussianDiffusion:
    """Sampling half of the reference's GaussianDiffusion (gaussian_diffusion.py:139)."""

    def __init__(self, conf):
        self.conf = conf
        self.model_mean_type = conf.model_mean_type
        self.model_var_type = conf.model_var_type
        self.rescale_timesteps = bool(getattr(conf, "rescale_timesteps", False))
        betas = np.array(conf.betas, dtype=np.float64)
        self.betas = betas
        assert len(betas.shape) == 1, "betas must be 1-D"
        assert (betas > 0).all() and (betas <= 1).all()
        self.num_timesteps = int(betas.shape[0])
        alphas = 1.0 - betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.alphas_cumprod_prev = np.append(1.0, self.alphas_cumprod[:-1])
        self.alphas_cumprod_next = np.append(self.alphas_cumprod[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - self.alphas_cumprod)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - self.alphas_cumprod)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / self.alphas_cumprod - 1)
        self.posterior_variance = betas * (1.0 - self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_log_variance_clipped = np.log(np.append(self.posterior_variance[1], self.posterior_variance[1:]))
        self.posterior_mean_coef1 = betas * np.sqrt(self.alphas_cumprod_prev) / (1.0 - self.alphas_cumprod)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - self.alphas_cumprod)
        if not hasattr(self, "timestep_map"):
            self.timestep_map = list(range(self.num_timesteps))
        # hook attributes the reference's eval code sets (gaussian_diffusion.py:238-241)
        self.data_transform_fn = None
        self.data_inv_transform_fn = None
        self.data_get_mean_fn = None
        self.log_trajectory_fn = None
        # engine options
        self.precision = capi.PRECISION_BF16X3
        self.max_batch = None          # default: the batch of the first call
        self.noise_tape = None         # test aid: (1 + num_steps, B, njoints, 1, nframes) on the device
        self.use_graph = True
        self.sample_offset = 0         # global index of local sample 0 (multi-GPU sharding)
        # per-step noise: "torch" = the exact stream torch.randn_like would produce on this device from torch's CUDA
        # generator (a reference GPU run with the same seed draws the same noise; the generator is advanced as the
        # reference loop would advance it); "engine" = the engine's own counter-based generator keyed by the global
        # sample index (sharding-independent; used by distributed.sharded_sample)
        self.rng = "torch"
        self.engine_seed = None  # rng="engine": explicit Philox key (sharded_sample broadcasts rank 0's); None = draw one

    # ------------------------------------------------------------------------------------------
    def q_sample(self, x_start, t, noise=None):
        """gaussian_diffusion.py:311-328 (host-side torch; the loop's own q_sample runs in the engine)."""
        if noise is None:
            noise = torch.randn_like(x_start)
        assert noise.shape == x_start.shape
        return (_extract_into_tensor(self.sqrt_alphas_cumprod, t, x_start.shape) * x_start +
                _extract_into_tensor(self.sqrt_one_minus_alphas_cumprod, t, x_start.shape) * noise)

    # ------------------------------------------------------------------------------------------
    def _check_supported(self, cond_fn, const_noise, randomize_class, model_kwargs):
        if const_noise:
            raise NotImplementedError()  # gaussian_diffusion.py:698-699, :1480-1481
        if randomize_class:
            raise NotImplementedError("randomize_class is a class-conditional feature the MDM path does not use")
        if _enum_name(self.model_mean_type) != "START_X" or _enum_name(self.model_var_type) != "FIXED_SMALL":
            raise NotImplementedError("the B200 engine implements START_X / FIXED_SMALL (utils/model_util.py:142-148)")
        if self.rescale_timesteps:
            raise NotImplementedError("rescale_timesteps=True is not used by the reference (model_util.py:130)")
        y = model_kwargs["y"]  # the reference indexes it unconditionally (gaussian_diffusion.py:1280)
        if "gmd" in y.keys():
            raise NotImplementedError("'gmd' selects p_sample_with_grad (classifier guidance): out of scope")
        if y.get("reconstruction_guidance", False):
            assert "stop_recguidance_at" in y.keys()  # utils/editing_util.py:329-330
            assert "inpainting_mask" in y.keys() and "inpainted_motion" in y.keys()
        return y

    def _run(self, sampler, model, shape, noise, cond_fn, model_kwargs, device, skip_timesteps, init_image, randomize_class,
             dump_steps, const_noise, eta, progressive=False):
        if model_kwargs is None:
            model_kwargs = {}
        y = self._check_supported(cond_fn, const_noise, randomize_class, model_kwargs)
        if sampler == capi.SAMPLER_DDPM:
            assert cond_fn is None, "only support the case where cond_fn is None"  # gaussian_diffusion.py:685
        elif cond_fn is not None:
            raise NotImplementedError("cond_fn (condition_score_with_grad) is out of scope")
        assert isinstance(shape, (tuple, list))
        inner, is_cfg = resolve_model(model)
        if device is None:
            device = next(model.parameters()).device
        device = torch.device(device)
        B = int(shape[0])
        eng = inner.engine_for(device, max_batch=max(B, self.max_batch or 0), precision=self.precision,
                                nframes=int(shape[-1]))
        eng.set_schedule(self.betas, self.timestep_map)

        # ---- conditioning ----
        cond_emb, text_scale, uncond = None, None, bool(y.get("uncond", False))
        cond_mode = getattr(inner, "cond_mode", "no_cond")
        if "action" in cond_mode:
            raise NotImplementedError("action conditioning is not on the HumanML3D path")
        if "text" in cond_mode:
            cond_emb = inner.encode_text(y["text"]).to(device=device, dtype=torch.float32)  # once per loop (mdm.py:249 does it per step)
        if is_cfg:
            assert cond_mode in ["text", "action"]  # cfg_sampler.py:27
            text_scale = y["text_scale"].to(device=device, dtype=torch.float32).reshape(-1)
        # ---- keyframe imputation (gaussian_diffusion.py:427-442, editing_util.py:336-346) ----
        imputate, stop_at, obs, mask, y_mask = False, 0, None, None, None
        guided_cfg = bool(y.get("reconstruction_guidance", False))
        if "imputate" in y.keys() and y["imputate"]:
            assert "stop_imputation_at" in y.keys()
            assert "inpainting_mask" in y.keys() and "inpainted_motion" in y.keys()
            # the reconstruction-guidance branch (:405-425) imputes whenever requires_imputation() holds and never reads
            # replacement_distribution; only the un-guided branch (:427-442) dispatches on it
            dist = "conditional" if guided_cfg else y["replacement_distribution"]
            if dist == "conditional":
                imputate, stop_at = True, int(y["stop_imputation_at"])
                obs = y["inpainted_motion"].to(device=device, dtype=torch.float32)
                mask = y["inpainting_mask"].to(device=device)
                assert obs.shape == mask.shape == tuple(shape)
                y_mask = y["mask"].to(device=device).reshape(B, -1)
            elif dist == "marginal":
                pass  # the reference's 'marginal' branch only calls the model (:437-439)
            else:
                raise NotImplementedError
        # ---- reconstruction guidance (gaussian_diffusion.py:405-425, editing_util.py:325-333) ----
        recon, stop_rg, coef = False, 0, None
        if y.get("reconstruction_guidance", False):
            recon, stop_rg = True, int(y["stop_recguidance_at"])
            if obs is None:
                obs = y["inpainted_motion"].to(device=device, dtype=torch.float32)
                mask = y["inpainting_mask"].to(device=device)
                assert obs.shape == mask.shape == tuple(shape)  # :414
                y_mask = y["mask"].to(device=device).reshape(B, -1)
            # w_r[t] * sqrt(alpha_bar_t) / 2, in fp32 exactly as the reference forms it (:418-422); the schedule is
            # indexed by the sampler step t like _extract_into_tensor(grad_ws, t, ...) does
            ws = get_gradient_schedule(y["gradient_schedule"], y["diffusion_steps"])
            tt = torch.arange(self.num_timesteps)
            w_r = torch.from_numpy(ws)[tt].float() * y["reconstruction_weight"]
            sab = torch.from_numpy(self.sqrt_alphas_cumprod)[tt].float()
            coef = (w_r * sab / 2).float().cpu().numpy()
        # ---- noise ----
        tape = self.noise_tape  # tape[0]: the initial randn(*shape) draw; tape[1 + k]: the k-th randn_like draw
        engine_rng = tape is None and self.rng != "torch"
        if noise is not None:
            x_T = noise.to(device=device, dtype=torch.float32)
        elif tape is not None:
            x_T = tape[0]
        elif engine_rng:
            # the engine draws x_T itself, keyed by (seed, sample_offset + b): rank r of a sharded run starts its
            # sample i from the x_T a single-GPU run gives global sample r*B/G + i (distributed.sharded_sample)
            x_T = None
        else:
            x_T = torch.randn(*shape, device=device)  # the reference's first draw (:1248)
        if tape is not None:
            tape = tape[1:]
        seed, rng_args = 0, {}
        if tape is None:
            n_draws = self.num_timesteps - skip_timesteps  # one randn_like per loop iteration (:696, :1407)
            rng_args = _torch_stream_args(device, int(np.prod(shape)), n_draws, lazy=progressive) if self.rng == "torch" else None
            if rng_args is None:
                if x_T is not None and noise is None:
                    x_T = None  # torch-compatible stream unavailable in this build: fall through to the engine generator
                # per-step noise comes from the engine's counter-based generator, keyed by `engine_seed` (set by
                # sharded_sample: identical on all ranks) or by a draw from torch's global CPU generator, so `fixseed`
                # still makes runs reproducible
                seed = self.engine_seed if self.engine_seed is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
                rng_args = {}
            else:
                seed = rng_args.pop("seed")
        if skip_timesteps and init_image is None:
            init_image = torch.zeros(tuple(shape), device=device, dtype=torch.float32)
        if init_image is not None:
            init_image = init_image.to(device=device, dtype=torch.float32)
        # keyframe INPUT conditioning (MDM_UNET.forward's obs_x0 / obs_mask): top-level model_kwargs, as
        # sample/conditional_synthesis.py:159-162 passes them; the transformer ignores them
        kf_obs, kf_mask = None, None
        if eng.arch == capi.ARCH_UNET and model_kwargs.get("obs_x0") is not None:
            kf_obs = model_kwargs["obs_x0"].to(device=device, dtype=torch.float32)
            kf_mask = model_kwargs["obs_mask"].to(device=device)
        common = dict(batch=B, sampler=sampler, eta=eta, cond_emb=cond_emb, uncond=uncond, cfg=is_cfg, text_scale=text_scale,
                      obs_x0=kf_obs, obs_mask=kf_mask,
                      y_mask=y_mask, imputate=imputate, stop_imputation_at=stop_at, inpainted_motion=obs,
                      inpainting_mask=mask, seed=seed, sample_offset=self.sample_offset, use_graph=self.use_graph,
                      recon_guidance=recon, stop_recguidance_at=stop_rg, recon_coef=coef, **rng_args)
        if not progressive:
            res = eng.sample(skip_timesteps=skip_timesteps, init_image=init_image, x_T=x_T,
                             noise_tape=None if tape is None else tape, want_pred_xstart=False, dump_steps=dump_steps, **common)
            return res
        return self._progressive(eng, x_T, init_image, skip_timesteps, tape, common)

    def _progressive(self, eng, x_T, init_image, skip_timesteps, tape, common):
        """Generator form: one native call per step (slower than the fused loop; kept for API parity)."""
        n = self.num_timesteps - skip_timesteps
        state = x_T
        common = dict(common)
        off0, inc = common.pop("aten_offset", 0), common.get("aten_increment", 0)
        for k in range(n):
            if common.get("rng_mode", capi.RNG_ENGINE) == capi.RNG_TORCH:
                common["aten_offset"] = off0 + k * inc  # each one-step call starts at its own draw of the stream
            common["use_graph"] = 2 if common.get("use_graph", True) else 0  # one native call per step, same step graph every time
            res = eng.sample(skip_timesteps=skip_timesteps + k, num_steps=1, resume=(k > 0), init_image=init_image if k == 0 else None,
                             x_T=state, noise_tape=None if tape is None else tape[k:], want_pred_xstart=True, **common)
            state = res["sample"]
            if common.get("rng_mode", capi.RNG_ENGINE) == capi.RNG_TORCH:
                # torch's generator moves one randn_like at a time, as the reference's loop moves it: a caller that
                # breaks out of the generator early leaves it where the reference would
                _advance_torch_generator(eng.device, inc)
            yield {"sample": res["sample"], "pred_xstart": res["pred_xstart"]}

    # ------------------------------------------------------------------------------------------
    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                      device=None, progress=False, skip_timesteps=0, init_image=None, randomize_class=False,
                      cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """gaussian_diffusion.py:1149-1214. Returns the final sample, or the list of pred_xstart at `dump_steps`."""
        res = self._run(capi.SAMPLER_DDPM, model, shape, noise, cond_fn, model_kwargs, device, skip_timesteps, init_image,
                        randomize_class, dump_steps, const_noise, 0.0)
        if dump_steps is not None:
            return res["dump"]
        return res["sample"]

    def p_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                  model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                                  randomize_class=False, cond_fn_with_grad=False, const_noise=False):
        """gaussian_diffusion.py:1217-1297.  Returns the step generator; the configuration is validated at the call (the
        reference, a generator function, defers the same checks to the first next())."""
        return self._run(capi.SAMPLER_DDPM, model, shape, noise, cond_fn, model_kwargs, device, skip_timesteps,
                             init_image, randomize_class, None, const_noise, 0.0, progressive=True)

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False):
        """gaussian_diffusion.py:1454-1512."""
        if const_noise == True:  # noqa: E712  (:1480-1481)
            raise NotImplementedError()
        res = self._run(capi.SAMPLER_DDIM, model, shape, noise, cond_fn, model_kwargs, device, skip_timesteps, init_image,
                        randomize_class, dump_steps, const_noise, eta)
        if dump_steps is not None:
            return res["dump"]
        return res["sample"]

    def ddim_sample_loop_progressive(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                                     model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0,
                                     init_image=None, randomize_class=False, cond_fn_with_grad=False):
        """gaussian_diffusion.py:1514-1587."""
        return self._run(capi.SAMPLER_DDIM, model, shape, noise, cond_fn, model_kwargs, device, skip_timesteps,
                             init_image, randomize_class, None, False, eta, progressive=True)


def space_timesteps(num_timesteps, section_counts):
    """respace.py:9-62."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            desired_count = int(section_counts[len("ddim"):])
            for i in range(1, num_timesteps):
                if len(range(0, num_timesteps, i)) == desired_count:
                    return set(range(0, num_timesteps, i))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(x) for x in section_counts.split(",")]
    size_per = num_timesteps // len(section_counts)
    extra = num_timesteps % len(section_counts)
    start_idx = 0
    all_steps = []
    for i, section_count in enumerate(section_counts):
        size = size_per + (1 if i < extra else 0)
        if size < section_count:
            raise ValueError(f"cannot divide section of {size} steps into {section_count}")
        frac_stride = 1 if section_count <= 1 else (size - 1) / (section_count - 1)
        cur_idx = 0.0
        taken_steps = []
        for _ in range(section_count):
            taken_steps.append(start_idx + round(cur_idx))
            cur_idx += frac_stride
        all_steps += taken_steps
        start_idx += size
    return set(all_steps)


class SpacedDiffusion(GaussianDiffusion):
    """respace.py:65-116: keeps `use_timesteps` of the base process; the engine applies timestep_map on the device."""

    def __init__(self, use_timesteps, conf):
        self.use_timesteps = set(use_timesteps)
        self.timestep_map = []
        self.original_num_steps = len(conf.betas)
        base = GaussianDiffusion.__new__(GaussianDiffusion)
        GaussianDiffusion.__init__(base, conf)
        last_alpha_cumprod = 1.0
        new_betas = []
        for i, alpha_cumprod in enumerate(base.alphas_cumprod):
            if i in self.use_timesteps:
                new_betas.append(1 - alpha_cumprod / last_alpha_cumprod)
                last_alpha_cumprod = alpha_cumprod
                self.timestep_map.append(i)
        new_conf = DiffusionConfig(betas=np.array(new_betas), model_mean_type=conf.model_mean_type,
                                   model_var_type=conf.model_var_type,
                                   rescale_timesteps=bool(getattr(conf, "rescale_timesteps", False)))
        super().__init__(new_conf)


def create_gaussian_diffusion(noise_schedule="cosine", steps=1000, use_ddim=False, timestep_respacing=None,
                              predict_xstart=True, sigma_small=True):
    """utils/model_util.py:122-165 (the sampling-relevant arguments)."""
    if timestep_respacing is None:
        timestep_respacing = "ddim100" if use_ddim else ""
    betas = get_named_beta_schedule(noise_schedule, steps, 1.)
    if not timestep_respacing:
        timestep_respacing = [steps]
    return SpacedDiffusion(
        use_timesteps=space_timesteps(steps, timestep_respacing),
        conf=DiffusionConfig(betas=betas,
                             model_mean_type=ModelMeanType.START_X if predict_xstart else ModelMeanType.EPSILON,
                             model_var_type=ModelVarType.FIXED_SMALL if sigma_small else ModelVarType.FIXED_LARGE,
                             rescale_timesteps=False))


def from_reference_diffusion(ref_diffusion) -> SpacedDiffusion:
    """Build the engine-backed sampler from a reference GaussianDiffusion/SpacedDiffusion instance."""
    d = SpacedDiffusion.__new__(SpacedDiffusion)
    d.timestep_map = list(getattr(ref_diffusion, "timestep_map", range(ref_diffusion.num_timesteps)))
    d.use_timesteps = set(d.timestep_map)
    d.original_num_steps = getattr(ref_diffusion, "original_num_steps", ref_diffusion.num_timesteps)
    GaussianDiffusion.__init__(d, DiffusionConfig(betas=np.array(ref_diffusion.betas, dtype=np.float64),
                                                  model_mean_type=ref_diffusion.model_mean_type,
                                                  model_var_type=ref_diffusion.model_var_type,
                                                  rescale_timesteps=ref_diffusion.rescale_timesteps))
    for hook in ("data_transform_fn", "data_inv_transform_fn", "data_get_mean_fn", "log_trajectory_fn"):
        setattr(d, hook, getattr(ref_diffusion, hook, None))
    return d


_warned_policy = False
_policy_ok = {}


def aten_policy(numel: int, num_sms: int, max_threads_per_sm: int) -> tuple:
    """(threads, philox offset increment) of ATen's CUDA `normal_` kernel for `numel` elements
    (aten/src/ATen/native/cuda/DistributionTemplates.h: calc_execution_policy, block 256, unroll 4): the grid is
    min(SMs * resident blocks per SM, ceil(numel / 256)) blocks, every thread makes ceil(numel / (threads * 4)) calls
    of curand_normal4 and each call advances the generator offset by 4."""
    blocks = min(num_sms * (max_threads_per_sm // 256), (numel + 255) // 256)
    threads = 256 * blocks
    return threads, ((numel - 1) // (threads * 4) + 1) * 4


def aten_launch_policy(numel: int, device) -> tuple:
    prop = torch.cuda.get_device_properties(device)
    return aten_policy(numel, prop.multi_processor_count, prop.max_threads_per_multi_processor)


def _cuda_rng_state(device) -> tuple:
    st = torch.cuda.get_rng_state(device)
    if st.numel() != 16:
        raise NotImplementedError("unexpected CUDA generator state layout")
    raw = bytes(st.tolist())
    return int.from_bytes(raw[:8], "little"), int.from_bytes(raw[8:], "little")


def _advance_torch_generator(device, n_offsets: int) -> None:
    seed, off = _cuda_rng_state(device)
    new_state = torch.tensor(list(seed.to_bytes(8, "little") + (off + n_offsets).to_bytes(8, "little")), dtype=torch.uint8)
    torch.cuda.set_rng_state(new_state, device)


def _torch_stream_args(device, numel: int, n_draws: int, lazy: bool = False):
    """Engine arguments that continue torch's CUDA generator stream for `n_draws` randn_like draws of `numel`
    elements, and advance torch's generator past them.  None (with one warning) if this torch build's launch policy
    is not the one `aten_launch_policy` models -- verified by drawing one element block and watching the offset."""
    global _warned_policy
    key = (str(device), numel)
    try:
        threads, inc = aten_launch_policy(numel, device)
        seed, off = _cuda_rng_state(device)
        if key not in _policy_ok:
            probe = torch.cuda.get_rng_state(device)
            torch.empty(numel, device=device).normal_()
            _policy_ok[key] = _cuda_rng_state(device) == (seed, off + inc)
            torch.cuda.set_rng_state(probe, device)
        ok = _policy_ok[key]
    except Exception:  # noqa: BLE001  (state layout / property names differ in this torch build)
        ok = False
    if not ok or off % 4:
        if not _warned_policy:
            warnings.warn("torch-compatible noise stream unavailable for this torch build; using the engine generator")
            _warned_policy = True
        return None
    if not lazy:  # the fused loop consumes all its draws inside one native call; the generators advance per yielded step
        _advance_torch_generator(device, n_draws * inc)
    return dict(seed=seed, rng_mode=capi.RNG_TORCH, aten_offset=off, aten_increment=inc, aten_threads=threads)


def _extract_into_tensor(arr, timesteps, broadcast_shape):
    """gaussian_diffusion.py:2215-2228."""
    res = torch.from_numpy(arr).to(device=timesteps.device)[timesteps].float()
    while len(res.shape) < len(broadcast_shape):
        res = res[..., None]
    return res.expand(broadcast_shape)
