"""Slice sharding: no data-path collective, one gather of finished slices.  world_size > 1 runs on the
gloo backend (CPU) here; on the GPU box the same code runs over NCCL (bench.py --gpus N)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusion_models_dsdiff_b200 import parallel as PAR


def test_shard_range_partitions_a_brats_volume():
    sizes = [PAR.shard_range(155, r, 8) for r in range(8)]           # 155 = 8*19 + 3
    assert [e - b for b, e in sizes] == [20, 20, 20, 19, 19, 19, 19, 19]
    assert sizes[0][0] == 0 and sizes[-1][1] == 155
    assert all(sizes[i][1] == sizes[i + 1][0] for i in range(7))      # contiguous, disjoint, covering


@pytest.mark.parametrize("n,world", [(0, 4), (3, 8), (64, 8), (1, 1), (7, 2)])
def test_shard_range_edge_cases(n, world):
    got = [PAR.shard_range(n, r, world) for r in range(world)]
    assert sum(e - b for b, e in got) == n
    assert max(e - b for b, e in got) - min(e - b for b, e in got) <= 1
    with pytest.raises(ValueError):
        PAR.shard_range(n, world, world)


def test_slice_noise_is_independent_of_the_batch_split():
    full = PAR.slice_noise(list(range(10)), (1, 8, 8), 2025, "cpu")
    a = PAR.slice_noise([0, 1, 2], (1, 8, 8), 2025, "cpu")
    b = PAR.slice_noise([7], (1, 8, 8), 2025, "cpu")
    assert torch.equal(full[:3], a) and torch.equal(full[7:8], b)
    assert not torch.equal(full[0], full[1])
    assert not torch.equal(PAR.slice_noise([0], (1, 8, 8), 2025, "cpu", stream_id=1), full[:1])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _finished_slice(sid: int) -> torch.Tensor:
    """CPU stand-in for one sampled slice: depends on the slice id only (the CUDA sampler cannot run here)."""
    return PAR.slice_noise([sid], (1, 16, 16), 99, "cpu")[0] * 0.5 + sid


def _worker(rank, world, port, n_total, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = PAR.shard_range(n_total, rank, world)
        local = torch.stack([_finished_slice(i) for i in range(b, e)]) if e > b else torch.empty((0, 1, 16, 16))
        full = PAR.gather_slices(local, n_total)
        if rank == 0:
            torch.save(full, out_path)
        else:
            assert full is None
        with pytest.raises(ValueError):                      # wrong local count is an error, not a silent truncation
            PAR.gather_slices(torch.zeros((e - b + 1, 1, 16, 16)), n_total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 8), (2, 7), (3, 155)])
def test_gather_slices_gloo_matches_single_process(world, n_total, tmp_path):
    out = str(tmp_path / "full.pt")
    mp.spawn(_worker, args=(world, _free_port(), n_total, out), nprocs=world, join=True)
    full = torch.load(out)
    ref = torch.stack([_finished_slice(i) for i in range(n_total)])
    assert full.shape == ref.shape and torch.equal(full, ref)   # result independent of the world size


def test_gather_slices_single_process_identity():
    x = torch.randn(5, 1, 4, 4)
    assert PAR.gather_slices(x, 5) is x
    with pytest.raises(ValueError):
        PAR.gather_slices(x, 6)


def test_pad_to_multiple_roundtrip():
    x = torch.randn(3, 2, 240, 240)
    p, pads = PAR.pad_to_multiple(x, 32)
    assert p.shape[-2:] == (256, 256) and pads == (8, 8, 8, 8)
    assert torch.equal(PAR.crop_pad(p, pads), x) and float(p[..., :8, :].abs().max()) == 0.0
    y = torch.randn(1, 1, 64, 96)
    assert PAR.pad_to_multiple(y, 32)[0] is y


def _fake_sampler(c, x_T):
    """Deterministic per-slice stand-in for a sampler (no cross-slice op, like the real network)."""
    return x_T * 0.25 + c[:, :1] * 2.0


def _volume_worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cond = torch.arange(13 * 3 * 24 * 24, dtype=torch.float32).reshape(13, 3, 24, 24) / 1e4
        full = PAR.sample_volume(_fake_sampler, cond, batch_size=3, base_seed=7, multiple=32)
        if rank == 0:
            torch.save(full, out_path)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sample_volume_is_independent_of_world_size_and_batch(world, tmp_path):
    out = str(tmp_path / "vol.pt")
    mp.spawn(_volume_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    cond = torch.arange(13 * 3 * 24 * 24, dtype=torch.float32).reshape(13, 3, 24, 24) / 1e4
    single = PAR.sample_volume(_fake_sampler, cond, batch_size=13, base_seed=7, multiple=32)
    assert single.shape == (13, 1, 24, 24)
    assert torch.equal(torch.load(out), single)


def _predict_worker(rank, world, port, tdir, odir):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from diffusion_models_dsdiff_b200 import volume_io as V
        cond = torch.arange(7 * 3 * 24 * 24, dtype=torch.float32).reshape(7, 3, 24, 24) / 1e4
        full = PAR.sample_volume(_fake_sampler, cond, batch_size=2, base_seed=11, multiple=32)
        assert (full is None) == (rank != 0)
        if rank == 0:            # rank 0 alone owns the 2D -> 3D assembly and the file (on_predict_end)
            asm = V.PredictionAssembler(["P7"])
            asm.add_paths([f"/data/test/P7/P7_{i}.h5" for i in range(7)], full)
            asm.finalize(tdir, odir)
    finally:
        dist.destroy_process_group()


def test_sharded_volume_to_nifti_end_to_end(tmp_path):
    """cfg4 glue on two gloo ranks: shard the slices, sample, gather, assemble, write <id>_pred.nii.gz from the template."""
    import numpy as np
    from test_volume_io import _write_template
    from diffusion_models_dsdiff_b200 import volume_io as V
    tdir, odir = str(tmp_path / "templates"), str(tmp_path / "out")
    _write_template(os.path.join(tdir, "P7", "T1CE.nii.gz"), np.zeros((7, 24, 24), np.float32))
    mp.spawn(_predict_worker, args=(2, _free_port(), tdir, odir), nprocs=2, join=True)
    cond = torch.arange(7 * 3 * 24 * 24, dtype=torch.float32).reshape(7, 3, 24, 24) / 1e4
    single = PAR.sample_volume(_fake_sampler, cond, batch_size=7, base_seed=11, multiple=32)
    vol, hdr = V.read_nifti(os.path.join(odir, "P7_pred.nii.gz"))
    assert hdr.shape == (7, 24, 24)
    assert np.array_equal(vol, single[:, 0].numpy())


# -------------------------------------------------------------------------------------------------
# cfg5: gradient arena layout + bucketed all-reduce of the DiT trainer (gloo, world size 2)
# -------------------------------------------------------------------------------------------------
def _dit_named_shapes():
    import diffusion_models_dsdiff_b200 as P
    with torch.device("meta"):
        net = P.DiT(input_size=16, patch_size=2, in_channels=4, hidden_size=64, depth=3, num_heads=2, learn_sigma=True)
    return [(k, p.numel()) for k, p in net.named_parameters() if p.requires_grad and not k.startswith("y_embedder.")]


def test_trainer_arena_layout_buckets_cover_the_arena_once():
    from diffusion_models_dsdiff_b200.dit_train import arena_layout
    named = _dit_named_shapes()
    slots, total, blocks, head, tail = arena_layout(named, 3)
    assert total % 8 == 0 and all(s.start % 8 == 0 for s in slots.values())
    order = [head] + blocks + [tail]
    assert order[0].start == 0 and order[-1].stop == total
    for a, b in zip(order, order[1:]):
        assert a.stop == b.start                                    # contiguous, non-overlapping, in backward-friendly order
    for k, sl in slots.items():                                      # every parameter lies inside exactly one bucket
        owner = [r for r in order if r.start <= sl.start and sl.stop <= r.stop]
        assert len(owner) == 1, k
        if k.startswith("blocks.1."):
            assert owner[0] is blocks[1]


def _bucket_worker(rank, world, port, q):
    import torch.distributed as dist
    from diffusion_models_dsdiff_b200.dit_train import arena_layout, reduce_bucket
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        slots, total, blocks, head, tail = arena_layout(_dit_named_shapes(), 3)
        g = torch.Generator().manual_seed(rank)
        flat = torch.randn(total, generator=g)
        mine = flat.clone()
        handles = []
        for r in reversed(blocks):                                   # the order the backward completes them
            reduce_bucket(flat, r, None, handles)
        reduce_bucket(flat, head, None, handles)
        reduce_bucket(flat, tail, None, handles)
        for h in handles:
            h.wait()
        other = torch.randn(total, generator=torch.Generator().manual_seed(1 - rank))
        q.put((rank, len(handles), float((flat - (mine + other)).abs().max())))
    finally:
        dist.destroy_process_group()


def test_trainer_bucketed_allreduce_sums_every_gradient_once():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for rank, n_handles, err in res:
        assert n_handles == 5 and err < 1e-6, (rank, n_handles, err)
