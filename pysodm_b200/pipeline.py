"""Chunked host<->device pipeline behind `sim()`: simulate on the GPU, deliver the reference's layout in HOST memory.

The reference returns, per state variable, an array `(draws, time, *state_dims)` in host memory
(`/root/reference/src/pySODM/models/base.py:969-1042` + concat along "draws" `:569-571`).  For N = 10^6
influenza trajectories that is 18 GB (int32) / 36 GB (float64) -- more than it is sensible to hold twice on the
device and far more PCIe time (~0.3-0.7 s) than kernel time.  The pipeline therefore works on chunks of
trajectories with two slots of device buffers and two streams:

    compute stream : H2D of the chunk's drawn parameters -> simulation kernel ([row][time][chunk], coalesced)
                     -> `sodm_to_draws_major` per state variable into a device staging buffer
    copy stream    : DMA of the staging buffers into the caller's PINNED host arrays (contiguous slices
                     [lo:hi] of each variable, because "draws" is the leading axis)

so that the D2H copy of chunk c overlaps the kernels of chunk c+1.  Per-trajectory RNG streams are keyed by the
global trajectory id, so the result does not depend on the chunk size.
"""
import numpy as np
import torch

from . import engine


class Variable:
    """One state variable in the flat state vector: rows [r0, r0 + n_dim)."""

    def __init__(self, name, r0, shape):
        self.name, self.r0, self.shape = name, int(r0), tuple(int(s) for s in shape)
        self.n_dim = int(np.prod(self.shape)) if self.shape else 1


def layout_from_shapes(state_shapes):
    out, r0 = [], 0
    for name, shape in state_shapes.items():
        v = Variable(name, r0, shape)
        out.append(v)
        r0 += v.n_dim
    return out, r0


def alloc_host_result(variables, n_traj, n_t, dtype, pin=True):
    """Pinned host arrays {name: tensor[n_traj, n_t, *shape]} that the pipeline fills."""
    res = {}
    for v in variables:
        t = torch.empty((n_traj, n_t) + v.shape, dtype=dtype, pin_memory=pin)
        res[v.name] = t
    return res


class HostPipeline:
    """Reusable two-slot pipeline.  `launch(lo, n, out_buf, stream)` must queue the simulation of trajectories
    [lo, lo+n) into out_buf[:, :, :n] ([n_row, n_t, chunk] device tensor) on `stream`."""

    def __init__(self, variables, n_row, n_t, chunk, sim_dtype, host_dtype, dev=None):
        self.dev = dev or engine.device()
        self.variables, self.n_row, self.n_t, self.chunk = variables, n_row, n_t, int(chunk)
        self.sim_dtype, self.host_dtype = sim_dtype, host_dtype
        self.sim_buf = [torch.empty((n_row, n_t, self.chunk), dtype=sim_dtype, device=self.dev) for _ in range(2)]
        self.stage = [{v.name: torch.empty((self.chunk, n_t, v.n_dim), dtype=host_dtype, device=self.dev)
                       for v in variables} for _ in range(2)]
        self.compute = torch.cuda.Stream(device=self.dev)
        self.copy = torch.cuda.Stream(device=self.dev)
        self.done_compute = [torch.cuda.Event() for _ in range(2)]
        self.done_copy = [torch.cuda.Event() for _ in range(2)]
        self.kernel_launches = 0

    def run(self, n_traj, launch, host_out):
        """Queues everything and returns; call `synchronize()` before touching host_out."""
        cur = torch.cuda.current_stream(self.dev)
        self.compute.wait_stream(cur)
        self.copy.wait_stream(cur)
        used = [False, False]
        for c, lo in enumerate(range(0, n_traj, self.chunk)):
            n = min(self.chunk, n_traj - lo)
            s = c & 1
            with torch.cuda.stream(self.compute):
                if used[s]:
                    self.compute.wait_event(self.done_copy[s])        # slot free again?
                launch(lo, n, self.sim_buf[s], self.compute)
                self.kernel_launches += 1
                for v in self.variables:
                    self._layout(s, v, n)
                    self.kernel_launches += 1
                self.done_compute[s].record(self.compute)
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(self.done_compute[s])
                for v in self.variables:
                    dst = host_out[v.name][lo:lo + n].view(n, self.n_t, v.n_dim)
                    dst.copy_(self.stage[s][v.name][:n], non_blocking=True)
                self.done_copy[s].record(self.copy)
            used[s] = True
        cur.wait_stream(self.copy)

    def _layout(self, s, v, n):
        # [row][time][chunk] -> staging [n][time][dim]; a shorter last chunk keeps the row stride `chunk`
        lib = engine._ffi.load()
        src, dst = self.sim_buf[s], self.stage[s][v.name]
        engine._ffi.check(lib.sodm_to_draws_major(src.data_ptr(), engine._TORCH_DT[src.dtype], int(src.stride(1)),
                                                  v.r0, v.n_dim, self.n_t, int(n), dst.data_ptr(),
                                                  engine._TORCH_DT[dst.dtype], engine._stream_ptr(self.compute)), lib)

    def synchronize(self):
        self.copy.synchronize()
        self.compute.synchronize()
