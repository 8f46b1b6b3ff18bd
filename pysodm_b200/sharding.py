"""Multi-GPU plumbing: contiguous trajectory / particle shards, one process per GPU, collectives only at the end.

The reference's only parallelism is a process pool over independent simulations
(`/root/reference/src/pySODM/models/base.py:559-562`, `optimization/pso.py:152-183`): units never talk to each
other.  On a B200 box the same units are sharded as contiguous ranges `[r*N/G, (r+1)*N/G)` over the ranks of a
`torch.distributed` job (backend "nccl" on GPUs; "gloo" in the CPU tests); every rank runs its range with Philox
streams keyed by the GLOBAL trajectory id, so the union of the shards is bit-identical to a single-GPU run, and the
only communication is the final gather of results (`lp[P]` is 8 bytes per particle; full trajectory tensors are
gathered only on request).
"""
import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def agree_on(value):
    """Rank 0's integer (a seed drawn from NumPy's global generator, say) on every rank."""
    rank, ws = world()
    if ws == 1:
        return int(value)
    box = [int(value)]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def shard_range(n_total, rank=None, world_size=None):
    """Contiguous [lo, hi) of `n_total` units owned by `rank`; the first n_total % G ranks get one extra unit."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_concat(local, n_total, axis=0):
    """Concatenate the ranks' shards along `axis` on every rank (shards may differ in length by one)."""
    rank, ws = world()
    if ws == 1:
        return local
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    origin = t.device
    if dist.get_backend() == "gloo" and t.is_cuda:            # gloo collectives run on host memory
        t = t.cpu()
    t = t.movedim(axis, 0).contiguous()
    sizes = [shard_range(n_total, r, ws)[1] - shard_range(n_total, r, ws)[0] for r in range(ws)]
    longest = max(sizes)
    if t.shape[0] < longest:                                  # pad to a common length: all_gather needs equal shapes
        pad = torch.zeros((longest - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad])
    parts = [torch.empty_like(t) for _ in range(ws)]
    dist.all_gather(parts, t)
    out = torch.cat([p[:s] for p, s in zip(parts, sizes)]).movedim(0, axis)
    return out.to(origin) if isinstance(local, torch.Tensor) else out.cpu().numpy()


def batch_sharded(objective, thetas, **kw):
    """`objective.batch` over a population sharded across the ranks; every rank returns the full lp[P]."""
    import numpy as np
    thetas = np.asarray(thetas)
    lo, hi = shard_range(thetas.shape[0])
    local = objective.batch(thetas[lo:hi], **kw) if hi > lo else np.zeros(0)
    rank, ws = world()
    if ws == 1:
        return local
    backend = dist.get_backend()
    t = torch.from_numpy(np.ascontiguousarray(local, dtype=np.float64))
    if backend == "nccl":
        t = t.cuda()
    return all_gather_concat(t, thetas.shape[0]).cpu().numpy()


def batch_device_sharded(objective, thetas, seed=None):
    """`objective.batch_device` for a device-resident population thetas[D, P] REPLICATED on every rank: each rank scores
    its contiguous range of columns (Philox streams keyed by one agreed seed and the global column index), the final
    all-gather of lp (8 bytes per member) is the only exchange; every rank returns the full lp[P] on its device."""
    rank, ws = world()
    if ws == 1:
        return objective.batch_device(thetas) if seed is None else objective.batch_device(thetas, seed=seed)
    if seed is None:                                          # a caller's seed must already be the same on every rank
        seed = agree_on(int(np.random.randint(0, 2 ** 32)) | (int(np.random.randint(0, 2 ** 31)) << 32))
    P = int(thetas.shape[1])
    lo, hi = shard_range(P)
    if hi > lo:
        local = objective.batch_device(thetas[:, lo:hi], seed=seed, member_id0=lo)
    else:
        local = torch.zeros(0, dtype=torch.float64, device=thetas.device)
    return all_gather_concat(local, P)


def same_on_every_rank(array):
    """Rank 0's copy of a host array on every rank (starting positions drawn from a rank's own generator)."""
    rank, ws = world()
    if ws == 1:
        return array
    box = [np.asarray(array)]
    dist.broadcast_object_list(box, src=0)
    return box[0]


# ---- reductions over draws of an ensemble sharded by trajectory --------------------------------------------------------
def _all_reduce(t, op):
    """In-place all-reduce of a (CUDA) tensor; gloo (CPU tests, single-GPU multi-process tests) goes through host memory."""
    if dist.get_backend() == "gloo" and t.is_cuda:
        h = t.cpu()
        dist.all_reduce(h, op=op)
        t.copy_(h)
    else:
        dist.all_reduce(t, op=op)
    return t


def sum_sharded(local_sum):
    """Sum over ALL ranks' draws from the per-rank sums (`engine.draws_sum`); counts stay exact (FP64 of int64 sums)."""
    rank, ws = world()
    return local_sum if ws == 1 else _all_reduce(local_sum.clone(), dist.ReduceOp.SUM)


def order_stats_sharded(ens_local, global_ranks):
    """The global_ranks[j]-th smallest value (0-based, over the union of all ranks' draws) of every cell of an ensemble
    whose trajectory axis is sharded over the ranks -> float64 [len(global_ranks), ...] on every rank.

    Radix select in phases (`sodm_draws_select_phase`): the data stays where it is; per pass one all-reduce of the
    per-cell histograms ([cells][ranks][2048] int32) -- the one real exchange on this path.  Single rank: the fused
    kernel (`engine.draws_order_stats`)."""
    from . import _ffi, engine
    rank, ws = world()
    if ws == 1:
        return engine.draws_order_stats(ens_local, global_ranks)
    outs = []
    global_ranks = [int(r) for r in global_ranks]
    for lo in range(0, len(global_ranks), _ffi.MAX_RANKS):
        search = engine.SelectPhases(ens_local, global_ranks[lo:lo + _ffi.MAX_RANKS])
        search.run(_ffi.SELECT_MINMAX)
        _all_reduce(search.lo, dist.ReduceOp.MIN)
        _all_reduce(search.hi, dist.ReduceOp.MAX)
        _all_reduce(search.nan_flag, dist.ReduceOp.MAX)
        for _ in range(-(-search.key_bits // 11)):
            search.run(_ffi.SELECT_HISTOGRAM)
            _all_reduce(search.hist, dist.ReduceOp.SUM)
            search.run(_ffi.SELECT_PICK)
        outs.append(search.result())
    return torch.cat(outs)


class _RawDeviceBuffer:
    """A device allocation of libsodm (sodm_peer_alloc / sodm_peer_open) as seen by `torch.as_tensor`."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


class PeerEnsemble:
    """The whole ensemble `[n_rows, n_out, n_total]` resident on EVERY GPU of the job, filled by the simulation kernels
    themselves: rank r simulates columns [lo_r, hi_r) and its kernel stores each output element into its own buffer and into
    the other ranks' buffers over NVLink (`sodm_jump_args.out_peers`).  This is the "final gather of outputs" of a sharded
    run without a collective: the peer stores are posted while the trajectories run, so the transfer hides behind the
    sampler instead of following it.  One node only (NVLink / NVSwitch peers).

    How the buffers are shared matters more than anything in the kernel.  `mapping='vmm'` (default under NCCL): CUDA
    virtual-memory allocations exported by file descriptor and mapped with 2 MB granularity -- torch's symmetric-memory
    allocator does the plumbing (`torch.distributed._symmetric_memory`), the kernel only sees raw pointers.
    `mapping='ipc'`: legacy CUDA IPC handles (`sodm_peer_alloc / sodm_peer_open`), which work between any two processes
    (also two processes sharing one GPU, as in the tests) but map the peer buffer with small pages: the kernel's stores
    scatter 128-byte rows over 18 GB per peer, and with three or more peers the translations thrash -- measured on 4 GPUs,
    250 000 trajectories each: kernel alone 29.6 ms; stores to 1 / 2 / 3 peers 30.5 / 31.3 / **204 ms** through IPC mappings,
    30.5 / 31.3 / **33.1 ms** through VMM mappings (profiles/r2_gather_probe_4gpu.txt); on 8 GPUs the IPC form took 873 ms."""

    def __init__(self, n_rows, n_out, n_total, dtype=torch.int32, dev=None, mapping="auto"):
        from . import _ffi, engine
        self.lib = _ffi.load()
        self.dev = dev or engine.device()
        self.rank, self.world = world()
        if self.world - 1 > _ffi.MAX_PEERS:
            raise ValueError(f"at most {_ffi.MAX_PEERS + 1} ranks, got {self.world}")
        self.shape = (int(n_rows), int(n_out), int(n_total))
        self.lo, self.hi = shard_range(n_total, self.rank, self.world)
        self.dtype = dtype
        if mapping == "auto":
            mapping = "vmm" if self.world > 1 and dist.get_backend() == "nccl" else "ipc"
        self.mapping = mapping
        self._own, self._mapped, self.peer_columns, self._handle = None, [], [], None      # peer_columns: column lo of every OTHER rank's buffer
        # every rank goes through the same collectives whatever fails locally, and all ranks raise together
        error = self._map_vmm() if mapping == "vmm" else self._map_ipc()
        if self.world > 1:
            errors = [None] * self.world
            dist.all_gather_object(errors, error)
            error = next((e for e in errors if e), None)
        if error is not None:
            self._release()
            raise RuntimeError(f"PeerEnsemble ({mapping}): {error}")
        self.local = self.full[:, :, self.lo:self.hi]          # what this rank's kernel writes (row stride n_total)

    def _map_vmm(self):
        itemsize = torch.empty(0, dtype=self.dtype).element_size()
        try:
            import torch.distributed._symmetric_memory as symm
            self.full = symm.empty(self.shape, dtype=self.dtype, device=self.dev)
            self._handle = symm.rendezvous(self.full, dist.group.WORLD)
            self._own = self.full.data_ptr()
            self._mapped = [int(p) for r, p in enumerate(self._handle.buffer_ptrs) if r != self.rank]
            self.peer_columns = [p + self.lo * itemsize for p in self._mapped]
        except Exception as exc:                                   # noqa: BLE001 -- reported on every rank together
            return f"{type(exc).__name__}: {exc}"
        return None

    def _map_ipc(self):
        import ctypes as C
        from . import _ffi
        itemsize = torch.empty(0, dtype=self.dtype).element_size()
        typestr = {torch.int32: "<i4", torch.float32: "<f4", torch.float64: "<f8"}[self.dtype]
        nbytes = itemsize * self.shape[0] * self.shape[1] * self.shape[2]
        ptr, handle, error = C.c_void_p(), C.create_string_buffer(_ffi.PEER_HANDLE_BYTES), None
        try:
            _ffi.check(self.lib.sodm_peer_alloc(nbytes, C.byref(ptr), handle), self.lib)
            self._own = ptr.value
        except _ffi.SodmError as exc:
            error = str(exc)
        handles = [(self.rank, torch.cuda.current_device(), handle.raw, error)]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, (self.rank, torch.cuda.current_device(), handle.raw, error))
        error = next((e for _, _, _, e in handles if e), None)
        if error is None:
            try:
                for r, device_index, raw, _ in handles:
                    if r == self.rank:
                        continue
                    _ffi.check(self.lib.sodm_peer_enable(int(device_index)), self.lib)
                    p = C.c_void_p()
                    _ffi.check(self.lib.sodm_peer_open(raw, C.byref(p)), self.lib)
                    self._mapped.append(p.value)
                    self.peer_columns.append(p.value + self.lo * itemsize)
                self.full = torch.as_tensor(_RawDeviceBuffer(self._own, self.shape, typestr), device=self.dev)
            except _ffi.SodmError as exc:
                error = str(exc)
        return error

    def push(self, col_lo, col_hi, stream=None):
        """The DMA form of the gather: copy columns [col_lo, col_hi) of this rank's shard (global column numbers) from its own
        copy of the ensemble into every peer's copy with the copy engines (`sodm_peer_copy2d`: one strided 2-D copy per peer, no SM
        involved).  Asynchronous on `stream`: queue it behind the kernel that produced the columns, on a side stream, and it
        overlaps the next kernel completely."""
        import ctypes as C
        from . import _ffi, engine
        itemsize = self.full.element_size()
        rows, pitch = self.shape[0] * self.shape[1], self.shape[2] * itemsize
        src = self._own + col_lo * itemsize
        sp = engine._stream_ptr(stream)
        for base in self._mapped:
            _ffi.check(self.lib.sodm_peer_copy2d(C.c_void_p(base + col_lo * itemsize), pitch, C.c_void_p(src), pitch,
                                                 (col_hi - col_lo) * itemsize, rows, sp), self.lib)

    def fence(self):
        """Every rank's stores have landed everywhere: local stream drained, then a barrier of the job."""
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()

    def _release(self):
        from . import _ffi
        if self.mapping == "vmm":
            self._mapped, self.peer_columns = [], []
            if self.world > 1:
                dist.barrier()                   # nobody still stores into this rank's buffer
            self.full = self.local = self._handle = None           # torch's allocator unmaps and frees
            self._own = None
            return
        for p in self._mapped:
            _ffi.check(self.lib.sodm_peer_close(p), self.lib)
        self._mapped = []
        if self.world > 1:
            dist.barrier()                       # nobody still maps this rank's buffer
        self.full = self.local = None
        if self._own:
            _ffi.check(self.lib.sodm_peer_free(self._own), self.lib)
            self._own = None

    def close(self):
        self.fence()
        self._release()
