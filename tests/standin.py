"""Stand-ins for the REFERENCE's own objects, built from stock torch.nn blocks, for boxes without /root/reference.

`StockMDM` has the module tree of the reference's `MDM` (model/mdm.py:99-160: `input_process.poseEmbedding`,
`sequence_pos_encoder.pe`, `embed_timestep.time_embed`, `seqTransEncoder = nn.TransformerEncoder(...)`,
`output_process.poseFinal`, optional `embed_text`), so its `state_dict()` carries exactly the reference's keys and its
`forward` IS stock PyTorch (nn.TransformerEncoder / nn.MultiheadAttention / F.gelu), not this repo's oracle and not
this repo's kernels.  `StockDiffusion` carries the attributes `condmdi_b200.accelerate()` reads from a reference
`SpacedDiffusion` (diffusion/respace.py:65-116).  Neither class is known to the package: the drop-in path duck-types.
"""
import enum

import numpy as np
import torch
import torch.nn as nn


class _PositionalEncoding(nn.Module):
    def __init__(self, d_model, max_len=5000):
        super().__init__()
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2).float() * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1))


class _TimestepEmbedder(nn.Module):
    def __init__(self, d, pos):
        super().__init__()
        self.sequence_pos_encoder = pos
        self.time_embed = nn.Sequential(nn.Linear(d, d), nn.SiLU(), nn.Linear(d, d))

    def forward(self, timesteps):
        return self.time_embed(self.sequence_pos_encoder.pe[timesteps]).permute(1, 0, 2)


class _Proc(nn.Module):
    def __init__(self, name, i, o):
        super().__init__()
        setattr(self, name, nn.Linear(i, o))


class StockMDM(nn.Module):
    """arch='trans_enc', data_rep='hml_vec' denoiser out of stock modules; eval-mode forward as mdm.py:239-306."""

    def __init__(self, njoints=263, d=512, ff=1024, layers=8, heads=4, text=False):
        super().__init__()
        self.njoints, self.nfeats, self.latent_dim, self.num_heads = njoints, 1, d, heads
        self.data_rep, self.dataset, self.arch = "hml_vec", "humanml", "trans_enc"
        self.cond_mode = "text" if text else "no_cond"
        self.cond_mask_prob = 0.1
        self.keyframe_conditioned = False
        self.rot2xyz, self.translation = None, True
        self.input_process = _Proc("poseEmbedding", njoints, d)
        self.sequence_pos_encoder = _PositionalEncoding(d)
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=heads, dim_feedforward=ff, dropout=0.1, activation="gelu")
        self.seqTransEncoder = nn.TransformerEncoder(layer, num_layers=layers, enable_nested_tensor=False)
        self.embed_timestep = _TimestepEmbedder(d, self.sequence_pos_encoder)
        if text:
            self.embed_text = nn.Linear(512, d)
        self.output_process = _Proc("poseFinal", d, njoints)
        self.text_emb = None

    def encode_text(self, raw_text):
        return self.text_emb

    def forward(self, x, timesteps, y=None, cond_val=None, cond_mask=None):
        y = y or {}
        bs, njoints, nfeats, nframes = x.shape
        emb = self.embed_timestep(timesteps)
        if "text" in self.cond_mode:
            enc = self.encode_text(y["text"])
            if y.get("uncond", False):
                enc = torch.zeros_like(enc)
            emb = emb + self.embed_text(enc)
        h = self.input_process.poseEmbedding(x.permute(3, 0, 1, 2).reshape(nframes, bs, njoints * nfeats))
        xseq = torch.cat((emb, h), dim=0) + self.sequence_pos_encoder.pe[:nframes + 1]
        out = self.seqTransEncoder(xseq)[1:]
        out = self.output_process.poseFinal(out).reshape(nframes, bs, njoints, nfeats)
        return out.permute(1, 2, 3, 0)


class ClassifierFreeSampleModel(nn.Module):
    """Same class name and attribute surface as model/cfg_sampler.py:5-35 (the package recognises the wrapper by name)."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        assert self.model.cond_mask_prob > 0
        for a in ("rot2xyz", "translation", "njoints", "nfeats", "data_rep", "cond_mode", "keyframe_conditioned"):
            setattr(self, a, getattr(model, a))

    def forward(self, x, timesteps, y=None, obs_x0=None, obs_mask=None, **kwargs):
        y_uncond = dict(y)
        y_uncond["uncond"] = True
        out, out_uncond = self.model(x, timesteps, y), self.model(x, timesteps, y_uncond)
        return out_uncond + (y["text_scale"].view(-1, 1, 1, 1) * (out - out_uncond))


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()


class StockDiffusion:
    """What `accelerate()` reads off a reference SpacedDiffusion: the respaced betas, the step map and the enums."""

    def __init__(self, betas, timestep_map, original_num_steps=1000):
        self.betas = np.asarray(betas, dtype=np.float64)
        self.num_timesteps = len(self.betas)
        self.timestep_map = list(timestep_map)
        self.original_num_steps = original_num_steps
        self.model_mean_type = ModelMeanType.START_X
        self.model_var_type = ModelVarType.FIXED_SMALL
        self.rescale_timesteps = False
        self.data_transform_fn = self.data_inv_transform_fn = self.data_get_mean_fn = self.log_trajectory_fn = None
