"""Multi-GPU plumbing: the path shards by patch.

Every height map carries its own 1-texel border (main.cpp:135-141) and depends only on its
104-byte Quad, so ranks take contiguous leaf ranges and never exchange anything while computing.
The only exchange is the gather of finished patches into one buffer on every rank (K4).  On GPUs
that is the C++ gather object behind the C-ABI (PatchGather below: NCCL + CUDA IPC + the fused
K2 kernel); `gather_patches` is the same exchange on torch.distributed for the CPU (gloo) tests
of the host-side logic."""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """Contiguous, balanced [lo, hi) of `n_units` patches for `rank`; the shards tile [0, n) in
    rank order so the gathered buffer is in the reference's emission order."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def gather_patches(local, n_units=None):
    """All ranks contribute their shard (first dimension = patches, possibly ragged across ranks)
    and receive the concatenation in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    all_counts = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = [int(c.item()) for c in all_counts]
    if n_units is not None and sum(all_counts) != n_units:
        raise RuntimeError(f"shards hold {sum(all_counts)} patches, expected {n_units}")
    if len(set(all_counts)) == 1:                              # equal shards: one all-gather in place
        out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    # ragged shards (they differ by at most one patch): pad to the largest, gather, trim
    cmax = max(all_counts)
    padded = torch.zeros((cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * cmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded)
    return torch.cat([out[r * cmax:r * cmax + c] for r, c in enumerate(all_counts)])


class PatchGather:
    """K4 through the C-ABI (planet_gpu_gather_*, csrc/k4_gather.cu): every rank's finished height
    maps in one buffer on every rank.  The C++ side owns everything -- NCCL communicator, CUDA IPC
    mappings of the peers' buffers, the GPU-to-GPU arrival / release flags; this class only hands
    the 128-byte NCCL unique id from rank 0 to the others over the existing process group (any
    transport would do: the C++ test driver uses a file) and wraps the calls.

        g = PatchGather(n_quads_total, dim)                 # collective
        g.height_maps(quads_of_my_shard, first_quad, max_depth, params)   # K2 + gather, stream-ordered
        ... K3 on g.local(first_quad, n) ...
        g.wait(release=True)                                 # the whole buffer is complete on this rank
        full = g.gathered()                                  # float32[n_quads_total, dim, dim] view
        g.close()                                            # collective
    """

    def __init__(self, n_quads, dim, n_buffers=2, device=None):
        import ctypes as C
        from . import lib, PlanetGpuError
        self.L, self.C = lib(), C
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_quads, self.dim, self.n_buffers = n_quads, dim, n_buffers
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        uid = None
        if self.world > 1:
            box = [None]
            if self.rank == 0:
                buf = C.create_string_buffer(128)
                rc = self.L.planet_gpu_gather_unique_id(buf)
                box[0] = bytes(buf.raw) if rc == 0 else self.L.planet_gpu_last_error().decode()
            dist.broadcast_object_list(box, src=0)
            if not isinstance(box[0], bytes):
                raise PlanetGpuError("planet_gpu_gather_unique_id: " + str(box[0]))
            uid = C.create_string_buffer(box[0], 128)
        self.bytes = n_quads * dim * dim * 4
        self.handle = self.L.planet_gpu_gather_create(uid, self.rank, self.world, self.bytes, n_buffers)
        if not self.handle:
            raise PlanetGpuError("planet_gpu_gather_create: " + self.L.planet_gpu_last_error().decode())

    def _check(self, rc):
        from . import _check
        _check(rc)

    def _view(self, which):
        """float32[n_quads, dim, dim] tensor over gathered buffer `which` (memory owned by the C++ side)."""
        ptr = self.L.planet_gpu_gather_buffer(self.handle, which)
        n = self.n_quads * self.dim * self.dim
        iface = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        holder = type("GatherView", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.device).view(self.n_quads, self.dim, self.dim)

    def gathered(self, which=None):
        return self._view(self.L.planet_gpu_gather_last_buffer(self.handle) if which is None else which)

    def local(self, first_quad, n, which=None):
        return self.gathered(which)[first_quad:first_quad + n]

    def height_maps(self, quads, first_quad, max_depth, params, stream=None):
        """K2 with the gather fused in for this rank's quads (int64[n, 13] device tensor)."""
        from . import _stream
        self._check(self.L.planet_gpu_gather_height_maps(self.handle, self.C.byref(params), quads.data_ptr(), quads.shape[0],
                                                         first_quad, self.dim, max_depth, _stream(stream)))

    def wait(self, release=True, stream=None):
        from . import _stream
        self._check(self.L.planet_gpu_gather_wait(self.handle, 1 if release else 0, _stream(stream)))

    def nccl(self, spans, which=0, stream=None):
        """The plain collective on data already in the local buffer: spans = [(first_quad, n_quads)] per rank."""
        from . import _stream
        per = self.dim * self.dim * 4
        off = (self.C.c_int64 * self.world)(*[lo * per for lo, _ in spans])
        size = (self.C.c_int64 * self.world)(*[n * per for _, n in spans])
        self._check(self.L.planet_gpu_gather_nccl(self.handle, which, off, size, _stream(stream)))

    def barrier(self, stream=None):
        from . import _stream
        self._check(self.L.planet_gpu_gather_barrier(self.handle, _stream(stream)))

    def check(self):
        self._check(self.L.planet_gpu_gather_error(self.handle))

    def close(self):
        if self.handle:
            self.L.planet_gpu_gather_destroy(self.handle)
            self.handle = None
