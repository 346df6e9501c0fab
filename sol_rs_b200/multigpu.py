"""Multi-GPU split of the path-tracing hot path — new work, the reference is single-GPU.

Frames are independent given the scene (per-pixel RNG seed = tea(pixel, frame), pathtrace.rgen:47), so the
scene + BVH are replicated, rank r of R renders frames f = first + r (mod R), accumulates a per-rank SUM of
frame colours (SOLB_ACCUM_SUM), and ONE collective — a float32 sum-reduce of the W x H x float4 buffers to
rank 0 over NCCL/NVLink — is the path's only exchange; rank 0 then resolves sum / count with the reference's
display transform.  Single frames are split by interleaved row bands and all-gathered.

The exchange itself lives behind the C ABI (libsolb: solb_comm_init / solb_reduce_accum / solb_allgather_rows,
csrc/comm.cu), on the ctx stream, so a Rust or C host has the same multi-GPU path; this module is the Python
binding.  The only thing a host has to provide is the bootstrap: getting the 128-byte communicator id from
rank 0 to the other ranks (`Communicator.from_torch_distributed` uses a torch.distributed broadcast for it,
`Communicator.from_id` takes bytes that travelled any other way)."""
import ctypes

from . import _native as N


def frames_for_rank(rank, world, n_frames, first=0):
    """Frames of [first, first + n_frames) owned by `rank`: f = first + rank (mod world)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(first + rank, first + n_frames, world))


def bands_for_rank(rank, world, height, band_rows):
    """Image rows owned by `rank` in the tile split: bands b = rank (mod world) of band_rows rows, clipped to the image."""
    if world < 1 or not (0 <= rank < world) or band_rows < 1:
        raise ValueError("bad rank/world/band_rows")
    rows = []
    for b in range(rank, (height + band_rows - 1) // band_rows, world):
        rows.extend(range(b * band_rows, min((b + 1) * band_rows, height)))
    return rows


def tile_rows_for_rank(rank, world, band_rows):
    """TraceBindings(tile_rows=...) of `rank`: (first row, rows per band, stride between its bands)."""
    return (rank * band_rows, band_rows, world * band_rows)


def _prefer_host_nccl():
    """libsolb binds the NCCL copy that is already in the process.  PyTorch bundles its own libnccl.so.2 and resolves it by
    soname when it is imported, so in a Python process torch must be imported BEFORE libsolb maps the system copy:
    otherwise torch would later be handed that (possibly older) copy and fail on a missing symbol."""
    try:
        import torch  # noqa: F401
    except ImportError:
        pass


def unique_id():
    """ncclGetUniqueId through the C ABI: 128 bytes rank 0 hands to every other rank."""
    _prefer_host_nccl()
    buf = (ctypes.c_uint8 * N.COMM_ID_BYTES)()
    N.check(N.lib().solb_comm_unique_id(buf))
    return bytes(buf)


class Communicator:
    """One NCCL communicator bound to a Context (solb_comm_init).  All methods are collective and asynchronous on the
    context's stream."""

    def __init__(self, context, comm_id, rank, world):
        if len(comm_id) != N.COMM_ID_BYTES:
            raise ValueError("communicator id must be %d bytes" % N.COMM_ID_BYTES)
        _prefer_host_nccl()
        self.context, self.rank, self.world = context, int(rank), int(world)
        buf = (ctypes.c_uint8 * N.COMM_ID_BYTES).from_buffer_copy(comm_id)
        N.check(N.lib().solb_comm_init(context.handle, buf, self.rank, self.world), context.handle)

    from_id = classmethod(lambda cls, context, comm_id, rank, world: cls(context, comm_id, rank, world))

    @classmethod
    def from_torch_distributed(cls, context, group=None):
        """Bootstrap over an initialised torch.distributed process group (any backend): rank 0's id is broadcast as a
        Python object.  torch.distributed is used for this one host-side exchange only."""
        import torch.distributed as dist

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(context, box[0], rank, world)

    def info(self):
        r, w, v = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        N.check(N.lib().solb_comm_info(self.context.handle, ctypes.byref(r), ctypes.byref(w), ctypes.byref(v)), self.context.handle)
        return {"rank": r.value, "world": w.value, "nccl_version": v.value}

    def reduce_accum(self, sum_target, root=0, accum_out=None, render=None):
        """Sum the per-rank SOLB_ACCUM_SUM targets onto `root` and resolve there (sum / count, gamma 2.2, rgba8)."""
        h = lambda t: t.handle if t is not None else None
        N.check(N.lib().solb_reduce_accum(self.context.handle, sum_target.handle, int(root), h(accum_out), h(render)),
                self.context.handle)

    def allgather_rows(self, target, band_rows):
        """Tile split: every rank traced tile_rows_for_rank(rank, world, band_rows); afterwards all hold the whole image."""
        N.check(N.lib().solb_allgather_rows(self.context.handle, target.handle, int(band_rows)), self.context.handle)

    def close(self):
        if self.context is not None:
            N.lib().solb_comm_destroy(self.context.handle)
            self.context = None


# ---- host-side restatement of the two exchanges on torch tensors: used by the gloo CPU tests of the sharding logic (the
#      GPU path above never goes through these) ----

def reduce_accum(tensor, dst=0, group=None):
    """Sum the per-rank accumulation buffers (xyz = sum of frame colours, w = frame count) onto `dst`."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(tensor, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return tensor


def allgather_band_rows(image, rank, world, band_rows, group=None):
    """image: [H, W, C] tensor whose rows bands_for_rank(rank, ...) are valid; returns the whole image on every rank
    (same packing as solb_allgather_rows: per-rank chunks of whole bands, the last band padded)."""
    import torch
    import torch.distributed as dist

    height = image.shape[0]
    n_bands = (height + band_rows - 1) // band_rows
    per_rank = (n_bands + world - 1) // world
    chunk = torch.zeros((per_rank * band_rows,) + tuple(image.shape[1:]), dtype=image.dtype, device=image.device)
    mine = bands_for_rank(rank, world, height, band_rows)
    for k, b in enumerate(range(rank, n_bands, world)):
        rows = list(range(b * band_rows, min((b + 1) * band_rows, height)))
        chunk[k * band_rows:k * band_rows + len(rows)] = image[rows]
    parts = [torch.empty_like(chunk) for _ in range(world)]
    dist.all_gather(parts, chunk, group=group)
    out = image.clone()
    for r in range(world):
        for k, b in enumerate(range(r, n_bands, world)):
            rows = list(range(b * band_rows, min((b + 1) * band_rows, height)))
            out[rows] = parts[r][k * band_rows:k * band_rows + len(rows)]
    assert len(mine) == sum(1 for _ in mine)
    return out
