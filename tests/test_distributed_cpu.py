"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: row-band sharding + the single
all-gather reassemble exactly the single-process frame; one-view-per-rank gather keeps rank order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from keypointnerf_b200 import distributed as D


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(y0, ny, width):
    """Deterministic stand-in for a rendered band: value depends only on the global pixel."""
    ys = torch.arange(y0, y0 + ny, dtype=torch.float32)[:, None]
    xs = torch.arange(width, dtype=torch.float32)[None, :]
    base = ys * 1000.0 + xs
    return torch.stack([base, base + 0.25, base + 0.5], 0)  # (3, ny, W)


def _worker(rank, world, port, height, width, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = D.init_from_env("gloo")
    assert (r, w) == (rank, world)
    y0, ny = D.row_shard(height, rank, world)
    band = _fake_render(y0, ny, width)
    full = D.gather_rows(band, height, rank, world)
    views = D.gather_views(torch.full((3, 4, 4), float(rank)), world)
    depth = D.gather_rows(band[0], height, rank, world)
    ok = torch.equal(full, _fake_render(0, height, width)) and torch.equal(depth, _fake_render(0, height, width)[0])
    ok = ok and all(float(views[i].mean()) == float(i) for i in range(world))

    class FakeMarcher:   # stands in for RayMarcher.render: the band of a deterministic frame, plus a per-ray debug plane
        def render(self, *, K, RT, znear, zfar, x0, y0, step, nx, ny, out_device, **kw):
            assert (x0, step, nx, out_device) == (0, 1, width, "cuda") and kw == {"S_c": 4}
            f = _fake_render(y0, ny, width)
            return {"tex_fg": f, "alpha": f[1], "contrib": torch.zeros(ny * nx, 4)}

    sharded = D.render_frame_row_sharded(FakeMarcher(), K=None, RT=None, znear=2.0, zfar=5.0, width=width, height=height, rank=rank,
                                         world=world, S_c=4)
    ok = ok and set(sharded) == {"tex_fg", "alpha"} and torch.equal(sharded["tex_fg"], _fake_render(0, height, width))
    ok = ok and torch.equal(sharded["alpha"], _fake_render(0, height, width)[1])

    class LatticeMarcher:   # the lattice phase (x0, y0, step, step_y) of the same deterministic frame
        def render(self, *, K, RT, znear, zfar, x0, y0, step, step_y, nx, ny, out_device, **kw):
            f = _fake_render(0, height, width)[:, y0::step_y, x0::step]
            assert f.shape[1:] == (ny, nx) and out_device == "cuda"
            return {"tex_fg": f.contiguous(), "alpha": f[1].contiguous(), "contrib": torch.zeros(ny * nx, 4)}

    lat = D.render_frame_lattice_sharded(LatticeMarcher(), K=None, RT=None, znear=2.0, zfar=5.0, width=width, height=height,
                                         rank=rank, world=world, S_c=4)
    ok = ok and set(lat) == {"tex_fg", "alpha"} and torch.equal(lat["tex_fg"], _fake_render(0, height, width))
    ok = ok and torch.equal(lat["alpha"], _fake_render(0, height, width)[1])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("height", [8, 7])
def test_row_shard_all_gather_world2(height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, height, 6, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_lattice_phases_cover_every_pixel_once():
    for world in (1, 2, 3, 4, 6, 8, 16):
        sy, sx = D.lattice_shape(world)
        assert sy * sx == world and sy <= sx
        img = torch.arange(3 * 24 * 48, dtype=torch.float32).reshape(3, 24, 48)
        shards = []
        for r in range(world):
            y0, x0, a, b = D.lattice_phase(r, world)
            assert (a, b) == (sy, sx)
            shards.append(img[:, y0::a, x0::b])
        assert torch.equal(D.interleave_lattice(torch.stack(shards), world), img)
        assert torch.equal(D.interleave_lattice(torch.stack([s[0] for s in shards]), world), img[0])


def test_row_shard_covers_every_row_once():
    for h in (1, 7, 512, 1024, 1023):
        for w in (1, 2, 3, 4, 8):
            rows = []
            for r in range(w):
                y0, ny = D.row_shard(h, r, w)
                rows += list(range(y0, y0 + ny))
            assert rows == list(range(h))
