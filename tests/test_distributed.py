"""Host logic of the multi-GPU path on CPU: scene sharding and the single
all-gather, world_size 2, gloo backend."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cld_b200 import make_scenes
from cld_b200.distributed import gather_results, pack_results, shard_batch, shard_scenes, unpack_results


def test_shard_scenes_covers_everything_once():
    for S in (1, 2, 7, 256, 4096):
        for G in (1, 2, 3, 8):
            spans = [shard_scenes(S, G, r) for r in range(G)]
            assert spans[0][0] == 0 and spans[-1][1] == S
            assert all(spans[i][1] == spans[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    torch.manual_seed(0)
    traj, off, coll = torch.randn(5, 52, 6), torch.rand(5, 52) > 0.5, torch.randint(0, 9, (5,)).float()
    t2, o2, c2 = unpack_results(pack_results(traj, off, coll), 52)
    assert torch.equal(t2, traj) and torch.equal(o2, off) and torch.equal(c2, coll)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, S, A, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        aux, batch = make_scenes(S, A, seed=5, dense=True)
        full = dict(batch, **aux)
        mine = shard_batch(full, A, world, rank)
        s0, s1 = shard_scenes(S, world, rank)
        assert mine["cond_feat"].shape[0] == (s1 - s0) * A
        assert torch.equal(mine["extent"], full["extent"][s0 * A:s1 * A])
        # stand-in for the sampler's outputs: a deterministic function of the global row index
        rows = torch.arange(s0 * A, s1 * A, dtype=torch.float32)
        traj = rows[:, None, None] + torch.arange(52 * 6, dtype=torch.float32).reshape(1, 52, 6) / 1000
        off = (rows[:, None].long() + torch.arange(52)[None]) % 3 == 0
        coll = rows % 5
        sizes = [(shard_scenes(S, world, r)[1] - shard_scenes(S, world, r)[0]) * A for r in range(world)]
        gt, go, gc = gather_results(traj, off, coll, rows_per_rank=sizes)
        all_rows = torch.arange(S * A, dtype=torch.float32)
        ok = (gt.shape == (S * A, 52, 6) and torch.equal(gt[:, 0, 0], all_rows) and torch.equal(gc, all_rows % 5)
              and torch.equal(go, (all_rows[:, None].long() + torch.arange(52)[None]) % 3 == 0))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _run(S, A):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, S, A, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_gather_world2_gloo_equal_shards():
    _run(4, 3)


def test_gather_world2_gloo_ragged_shards():
    _run(5, 2)


def test_shard_batch_carries_the_context_encoder_inputs():
    """The raster and the history tensors of the context encoder (row a14) shard by scene like every other per-agent entry."""
    S, A = 5, 3
    B = S * A
    full = {"image": torch.arange(B).float().view(B, 1, 1, 1).expand(B, 2, 4, 4).contiguous(),
            "history_positions": torch.arange(B).float().view(B, 1, 1).expand(B, 31, 2).contiguous(),
            "history_yaws": torch.zeros(B, 31, 1), "curr_speed": torch.arange(B).float(), "scene_index": torch.arange(S).repeat_interleave(A),
            "map_names": ["m"] * B}
    seen = []
    for rank in range(2):
        mine = shard_batch(full, A, 2, rank)
        s0, s1 = shard_scenes(S, 2, rank)
        assert mine["image"].shape[0] == (s1 - s0) * A and mine["map_names"] == full["map_names"]
        assert torch.equal(mine["image"][:, 0, 0, 0], mine["curr_speed"]) and torch.equal(mine["history_positions"][:, 0, 0], mine["curr_speed"])
        seen.append(mine["curr_speed"])
    assert torch.equal(torch.cat(seen), full["curr_speed"])
