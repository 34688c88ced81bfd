"""CPU, world_size 2 over gloo: the batch-sharding + all-gather logic of condmdi_b200.distributed
(the N>1 path of bench.py) with a stand-in sampler, so no GPU is needed."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from condmdi_b200.distributed import shard_model_kwargs, sharded_sample


class StandInDiffusion:
    """Returns a value per sample that encodes everything sharding must get right."""
    sample_offset = 0
    noise_tape = None

    def p_sample_loop(self, model, shape, noise=None, model_kwargs=None, **kw):
        B = shape[0]
        self.seen_seed = getattr(self, "engine_seed", None)
        y = model_kwargs["y"]
        assert len(y["text"]) == B and y["text_scale"].shape[0] == B and y["mask"].shape[0] == B
        assert y["imputate"] == 1  # non-batched entries pass through
        out = torch.zeros(shape)
        for b in range(B):
            out[b] = (self.sample_offset + b) + 1000.0 * y["text_scale"][b] + float(len(y["text"][b])) * 1e-3 \
                + (noise[b].sum() if noise is not None else 0.0) + (self.noise_tape[:, b].sum() if self.noise_tape is not None else 0.0)
        return out


def _expected(B, shape, text, scale, noise, tape):
    out = torch.zeros((B,) + shape)
    for b in range(B):
        out[b] = b + 1000.0 * scale[b] + len(text[b]) * 1e-3 + noise[b].sum() + tape[:, b].sum()
    return out


def _worker(rank, world, port, B):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        shape = (3, 1, 5)
        text = ["x" * (i + 1) for i in range(B)]
        scale = torch.arange(B, dtype=torch.float32) * 0.5
        noise = torch.randn((B,) + shape, generator=g)
        tape = torch.randn((4, B) + shape, generator=g)
        d = StandInDiffusion()
        d.noise_tape = tape
        kw = {"y": {"text": text, "text_scale": scale, "mask": torch.ones(B, 1, 1, 5, dtype=torch.bool), "imputate": 1}}
        out = sharded_sample(d, None, (B,) + shape, model_kwargs=kw, noise=noise)
        assert out.shape == (B,) + shape
        assert torch.allclose(out, _expected(B, shape, text, scale, noise, tape), atol=1e-5)
        assert d.sample_offset == 0 and d.noise_tape is tape  # restored
        # every rank ran with the SAME engine noise key (rank 0's draw), although the ranks' own generators differ
        torch.manual_seed(100 + rank)
        sharded_sample(d, None, (B,) + shape, model_kwargs=kw, noise=noise)
        seeds = [None] * world
        dist.all_gather_object(seeds, d.seen_seed)
        assert seeds[0] is not None and all(s_ == seeds[0] for s_ in seeds), seeds
        assert getattr(d, "engine_seed", None) is None  # restored
        with pytest.raises(ValueError):
            sharded_sample(d, None, (B + 1,) + shape, model_kwargs=kw)
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(180)
def test_sharded_sample_world2_gloo():
    mp.spawn(_worker, args=(2, _free_port(), 6), nprocs=2, join=True)


def test_shard_model_kwargs_slices_only_batched_entries():
    B = 4
    kw = {"y": {"text": list("abcd"), "lengths": torch.arange(B), "scalar": 3, "vec5": torch.zeros(5)}, "obs_x0": torch.zeros(B, 2)}
    s = shard_model_kwargs(kw, 2, 4, B)
    assert s["y"]["text"] == ["c", "d"] and s["y"]["lengths"].tolist() == [2, 3] and s["y"]["scalar"] == 3
    assert s["y"]["vec5"].shape == (5,) and s["obs_x0"].shape == (2, 2)


def test_single_process_is_identity():
    d = StandInDiffusion()
    kw = {"y": {"text": ["a", "bb"], "text_scale": torch.tensor([1.0, 2.0]), "mask": torch.ones(2, 1, 1, 5), "imputate": 1}}
    out = sharded_sample(d, None, (2, 3, 1, 5), model_kwargs=kw)
    assert out.shape == (2, 3, 1, 5) and float(out[1, 0, 0, 0]) == pytest.approx(1 + 2000 + 2e-3)


# ------------------------------------------------------------------------------------------------
# evaluation-loop orchestration (condmdi_b200.eval_loop): job list -> units -> ranks -> gathered samples
# ------------------------------------------------------------------------------------------------
class StandInEvalDiffusion:
    rng, sample_offset, engine_seed = "torch", 0, None

    def p_sample_loop(self, model, shape, model_kwargs=None, clip_denoised=True, **kw):
        assert clip_denoised is False and self.rng == "engine"
        y = model_kwargs["y"]
        B = shape[0]
        assert len(y["text"]) == B and y["text_scale"].shape[0] == B and y["inpainted_motion"].shape[0] == B
        out = torch.zeros(shape)
        for b in range(B):  # a value only the right (seed, global motion index, per-sample inputs) combination produces
            out[b] = self.engine_seed * 1e-3 + (self.sample_offset + b) + 100.0 * y["text_scale"][b] + len(y["text"][b]) * 0.01 \
                + y["inpainted_motion"][b].sum()
        return out


def _eval_jobs():
    from condmdi_b200.eval_loop import build_jobs
    g = torch.Generator().manual_seed(5)
    batches = []
    for i in range(5):
        bs = 3 if i != 3 else 2   # one ragged batch (the last dataloader batch of an evaluation set)
        y = {"text": ["w" * (i + k + 1) for k in range(bs)], "text_scale": torch.rand(bs, generator=g), "imputate": 1,
             "inpainted_motion": torch.randn(bs, 4, 1, 6, generator=g), "stop_imputation_at": 0 if i < 4 else 1}
        batches.append(((bs, 4, 1, 6), {"y": y}))
    return build_jobs(batches, seed=10, mm_idxs=[1], mm_num_repeats=3)


def _eval_worker(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from condmdi_b200.eval_loop import run_eval_jobs
        jobs = _eval_jobs()
        d = StandInEvalDiffusion()
        got = run_eval_jobs(d, None, jobs, seed=7, merge=2)
        want = _eval_expected(jobs, 7)
        assert sorted(got) == list(range(len(jobs)))
        for i in got:
            assert torch.allclose(got[i], want[i], atol=1e-5), i
        assert d.rng == "torch" and d.sample_offset == 0 and d.engine_seed is None  # restored
    finally:
        dist.destroy_process_group()


def _eval_expected(jobs, seed):
    want, n = {}, 0
    for i, j in enumerate(jobs):
        y = j.model_kwargs["y"]
        out = torch.zeros(j.shape)
        for b in range(j.shape[0]):
            out[b] = seed * 1e-3 + (n + b) + 100.0 * y["text_scale"][b] + len(y["text"][b]) * 0.01 + y["inpainted_motion"][b].sum()
        want[i] = out
        n += j.shape[0]
    return want


def test_eval_loop_plan_merge_and_single_process():
    from condmdi_b200.eval_loop import eval_seed_number, plan, run_eval_jobs
    jobs = _eval_jobs()
    assert len(jobs) == 7 and [(j.batch_index, j.repeat) for j in jobs][:4] == [(0, 0), (1, 0), (1, 1), (1, 2)]
    assert jobs[2].seed_number == eval_seed_number(10, 1, 1) == 1_000_101       # comp_v6_model_dataset_condmdi.py:293
    units = plan(jobs, world=2, merge=2)
    # job 6 (batch 4) differs in a non-batched entry (stop_imputation_at) and may not be merged with job 5
    assert units == [[0, 1], [2, 3], [4, 5], [6]]
    want = _eval_expected(jobs, 3)
    for merge in (1, 2, 4):
        got = run_eval_jobs(StandInEvalDiffusion(), None, jobs, seed=3, merge=merge)
        assert all(torch.allclose(got[i], want[i], atol=1e-5) for i in range(len(jobs))), merge  # independent of the merge width


def test_eval_loop_two_ranks_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_eval_worker, args=(2, port), nprocs=2, join=True)
