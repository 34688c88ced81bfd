"""Putting the engine under the reference's own objects (the drop-in boundary).

    import condmdi_b200
    model, diffusion = create_model_and_diffusion(args, data)      # reference code, unchanged
    ...
    diffusion = condmdi_b200.accelerate(diffusion)                 # <- the one added line
    sample = diffusion.p_sample_loop(model, (B, 263, 1, 196), model_kwargs=model_kwargs, ...)

`accelerate` returns an engine-backed sampler exposing the same methods/attributes; `install` patches the
reference object in place instead (for call sites that keep their own reference to it).
"""
from __future__ import annotations

import types

from .diffusion import GaussianDiffusion, from_reference_diffusion

_LOOPS = ("p_sample_loop", "p_sample_loop_progressive", "ddim_sample_loop", "ddim_sample_loop_progressive")


def accelerate(ref_diffusion):
    """Engine-backed sampler equivalent to a reference GaussianDiffusion / SpacedDiffusion instance."""
    if isinstance(ref_diffusion, GaussianDiffusion):
        return ref_diffusion
    return from_reference_diffusion(ref_diffusion)


def install(ref_diffusion, fallback_to_reference: bool = False):
    """Replace the four sampling loops of a reference diffusion object by the engine's, in place.

    Configurations the engine does not implement raise NotImplementedError.  With fallback_to_reference=True those
    (and only those) are forwarded to the reference's original PyTorch loop instead -- an explicit opt-in for
    scripts that mix accelerated and non-accelerated features, never a silent CPU/eager path.
    """
    fast = from_reference_diffusion(ref_diffusion)
    ref_diffusion._condmdi_b200 = fast
    for name in _LOOPS:
        original = getattr(ref_diffusion, name)

        def make(name=name, original=original):
            def loop(self, *args, **kwargs):
                try:
                    # (the *_progressive loops validate their configuration when CALLED and return the step
                    # generator, so an unsupported configuration surfaces inside this try block too)
                    return getattr(fast, name)(*args, **kwargs)
                except NotImplementedError:
                    if fallback_to_reference:
                        return original(*args, **kwargs)
                    raise
            return loop

        setattr(ref_diffusion, name, types.MethodType(make(), ref_diffusion))
    return ref_diffusion
