"""world_size-2 gloo test of the pair-sharding host logic (lightglue_b200/sharding.py) on CPU.
The matcher itself needs a GPU; here a stand-in callable produces deterministic per-pair outputs so
that shard boundaries, ragged shards and the gather order are checked without any CUDA compute."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lightglue_b200.sharding import gather_matches, match_sharded, shard_range


def test_shard_range_partitions_everything():
    for total in (0, 1, 7, 32, 1024, 1027):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


class FakeMatcher(torch.nn.Module):
    """matches0[p, i] = (sum of the pair's keypoint x-coords) % 97 + i  -- depends only on the pair's data."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    def forward(self, data):
        k0, k1 = data["image0"]["keypoints"], data["image1"]["keypoints"]
        tag0 = (k0[..., 0].sum(-1).round().long() % 97)[:, None]
        tag1 = (k1[..., 0].sum(-1).round().long() % 97)[:, None]
        m, n = k0.shape[1], k1.shape[1]
        return {
            "matches0": tag0 + torch.arange(m)[None], "matches1": tag1 + torch.arange(n)[None],
            "matching_scores0": (tag0 + torch.arange(m)[None]).float() / 7, "matching_scores1": tag1.float().expand(-1, n) / 3,
        }


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(11)
        data = {
            "image0": {"keypoints": torch.rand(total, 5, 2, generator=g) * 100},
            "image1": {"keypoints": torch.rand(total, 6, 2, generator=g) * 100},
        }
        fm = FakeMatcher()
        got = match_sharded(fm, data, batch=3)
        want = fm(data)
        ok = all(torch.equal(got[k], want[k]) for k in want)
        ok = ok and got["matches0"].dtype == torch.int64 and got["matches0"].shape == (total, 5)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_match_sharded_world2_equals_single_process(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
