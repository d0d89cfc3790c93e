"""Multi-GPU partitioning of the ray-march path: one process per GPU, scene replicated, rays sharded.

The reference has no multi-GPU rendering; rays are independent, so the only
exchange step is ONE all-gather of the rendered shards (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  Partitions:
  * ``lattice_shape`` / ``gather_lattice`` / ``render_frame_lattice_sharded``: one frame split into the
    stride-lattice phases the reference itself renders a frame in (``src/model.py:916-923``), one
    phase per rank: every phase sees the whole image, so the ranks are load-balanced by construction
    (BASELINE config 4);
  * ``row_shard`` / ``gather_rows`` / ``render_frame_row_sharded``: contiguous row bands (frames whose
    size the lattice does not divide);
  * ``gather_views``: one novel view per rank (BASELINE config 5, render_dynamic.py-style sweep).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def row_shard(height: int, rank: int, world: int):
    """Contiguous row band [y0, y0+ny) of rank ``rank``; heights that do not divide evenly give the
    first ``height % world`` ranks one extra row."""
    base, rem = divmod(height, world)
    ny = base + (1 if rank < rem else 0)
    y0 = rank * base + min(rank, rem)
    return y0, ny


def _all_gather_stack(t: torch.Tensor, world: int) -> torch.Tensor:
    """One all-gather; returns (world, *t.shape).  The output is laid out as a dim-0 concatenation,
    which both NCCL and gloo accept for ``all_gather_into_tensor``."""
    t = t.contiguous()
    if t.dim() == 0:
        t = t.reshape(1)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.view(world, *t.shape)


def gather_rows(shard: torch.Tensor, height: int, rank: int, world: int) -> torch.Tensor:
    """All-gather row bands of a planar image (..., ny, W) into (..., height, W) on every rank."""
    if world == 1:
        return shard
    base, rem = divmod(height, world)
    if rem == 0:
        out = _all_gather_stack(shard, world)
        # (world, ..., ny, W) -> (..., world*ny, W)
        nd = shard.dim()
        perm = list(range(1, nd - 1)) + [0, nd - 1, nd]
        return out.permute(*perm).reshape(*shard.shape[:-2], height, shard.shape[-1])
    # ragged bands: pad to the tallest band, gather, then trim
    ny_max = base + 1
    pad = torch.zeros(*shard.shape[:-2], ny_max, shard.shape[-1], dtype=shard.dtype, device=shard.device)
    pad[..., : shard.shape[-2], :] = shard
    out = _all_gather_stack(pad, world)
    bands = [out[r][..., : row_shard(height, r, world)[1], :] for r in range(world)]
    return torch.cat(bands, dim=-2)


def lattice_shape(world: int):
    """(sy, sx) with sy * sx == world, as square as possible, sx >= sy: rank r renders the pixels
    (x, y) with x % sx == r % sx and y % sy == r // sx."""
    sy = 1
    for d in range(1, int(world ** 0.5) + 1):
        if world % d == 0:
            sy = d
    return sy, world // sy


def lattice_phase(rank: int, world: int):
    """(y0, x0, step_y, step_x) of rank ``rank``'s lattice phase."""
    sy, sx = lattice_shape(world)
    return rank // sx, rank % sx, sy, sx


def interleave_lattice(stack: torch.Tensor, world: int) -> torch.Tensor:
    """(world, ..., h, w) lattice phases -> (..., h*sy, w*sx): the inverse of the partition (the reference's
    ``pixel_shuffle`` assembly, ``src/model.py:935-938``)."""
    sy, sx = lattice_shape(world)
    lead = tuple(stack.shape[1:-2])
    h, w = stack.shape[-2:]
    x = stack.reshape(sy, sx, *lead, h, w)
    nd = len(lead)
    perm = [2 + i for i in range(nd)] + [nd + 2, 0, nd + 3, 1]   # (..., h, sy, w, sx)
    return x.permute(*perm).reshape(*lead, h * sy, w * sx)


def gather_lattice(shard: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather the ranks' lattice phases (..., H/sy, W/sx) into the full (..., H, W) image on every rank."""
    if world == 1:
        return shard
    return interleave_lattice(_all_gather_stack(shard, world), world)


def render_frame_lattice_sharded(marcher, *, K, RT, znear, zfar, width: int, height: int, rank: int, world: int,
                                 **render_kw) -> dict:
    """BASELINE config 4: ONE target frame, rank r renders lattice phase r (scene already bound on every rank), then ONE
    all-gather per output plane and a local un-permute.  Bit-identical to the single-GPU frame: rays are independent."""
    y0, x0, sy, sx = lattice_phase(rank, world)
    if height % sy or width % sx:
        return render_frame_row_sharded(marcher, K=K, RT=RT, znear=znear, zfar=zfar, width=width, height=height,
                                        rank=rank, world=world, **render_kw)
    res = marcher.render(K=K, RT=RT, znear=znear, zfar=zfar, x0=x0, y0=y0, step=sx, step_y=sy, nx=width // sx,
                         ny=height // sy, out_device="cuda", **render_kw)
    return {k: gather_lattice(v, world) for k, v in res.items() if v.dim() >= 2 and v.shape[-2] == height // sy
            and v.shape[-1] == width // sx}


def gather_views(frame: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather one rendered frame per rank into (world, ...) on every rank."""
    if world == 1:
        return frame[None]
    return _all_gather_stack(frame, world)


def render_frame_row_sharded(marcher, *, K, RT, znear, zfar, width: int, height: int, rank: int, world: int, **render_kw) -> dict:
    """ONE target frame split into contiguous row bands, one band per rank (the scene must already be bound
    to every rank's ``RayMarcher``), then one all-gather per output plane; not load-balanced (the top and bottom bands are
    mostly empty space): ``render_frame_lattice_sharded`` is the partition BASELINE config 4 uses.  ``render_kw`` goes to ``RayMarcher.render``
    (``S_c``, ``S_f``, ``fine``, ``engine``, ``ert_eps``).  Returns full-frame tensors on every rank, bit-identical to the
    single-GPU frame because rays are independent (``tests/test_gpu_parity.py`` checks strided/offset passes against the frame)."""
    y0, ny = row_shard(height, rank, world)
    res = marcher.render(K=K, RT=RT, znear=znear, zfar=zfar, x0=0, y0=y0, step=1, nx=width, ny=ny, out_device="cuda", **render_kw)
    return {k: gather_rows(v, height, rank, world) for k, v in res.items() if v.dim() >= 2 and v.shape[-2] == ny}
