"""Strong-scaling gather probe (torchrun, one rank per GPU): the C3 step with 10^6 trajectories in total, every GPU ending with
the whole ensemble -- (a) kernel alone, (b) in-kernel peer stores to k = 1 .. G-1 peers, (c) DMA push of 1/2/4/8 chunks on a
side stream overlapped with the next chunk's kernel.   torchrun --nproc-per-node G tools/fused_gather_probe.py [n_total]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
from pysodm_b200 import engine, sharding, tutorial_data as td

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", rank=rank, world_size=world)
dev = engine.device(torch.cuda.current_device())
n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
lo, hi = sharding.shard_range(n_total, rank, world)
n = hi - lo
y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
idx = torch.from_numpy(td.influenza_schedule()).to(dev)
draws = torch.from_numpy(bench.synthetic_draws(n_total, 1718)[:, lo:hi].copy()).to(dev)
status = torch.zeros(n, dtype=torch.int32, device=dev)
backend = os.environ.get("SODM_PROBE_BACKEND", "vmm")          # 'vmm' (CUDA virtual-memory mappings) or 'ipc' (legacy CUDA IPC)
ens = sharding.PeerEnsemble(28, 162, n_total, torch.int32, dev, mapping=backend)
flag = torch.zeros(1, dtype=torch.int32, device=dev)
side = torch.cuda.Stream(device=dev)
cur = torch.cuda.current_stream(dev)


def run(seed, c_lo, c_hi, peers):
    engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), c_hi - c_lo, 0, 162, 0.5, var=(bench.VAR_INDEX, draws[:, c_lo:c_hi].contiguous()),
                    tdpf=(tab, idx), seed=seed, traj_id0=lo + c_lo, out_dtype=torch.int32, out=ens.full[:, :, lo + c_lo:lo + c_hi],
                    status=status[c_lo:c_hi], dev=dev, out_peers=[p + c_lo * 4 for p in peers] if peers else None)


def timed(step, reps=4):
    for w in range(2):
        step(10 + w)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for k in range(reps):
        step(100 + k)
    b.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def say(*a):
    if rank == 0:
        print(*a, flush=True)


say(f"{world} GPUs, {n_total} trajectories in total, {n} per GPU, peer mappings: {backend}")
say(f"kernel alone: {timed(lambda s: run(s, 0, n, None)):.2f} ms")
for k in range(1, world):
    def step(s, k=k):
        run(s, 0, n, ens.peer_columns[:k])
        dist.all_reduce(flag)
    say(f"in-kernel stores to {k} peer(s): {timed(step):.2f} ms")
for chunks in (4,):
    m = n // chunks
    def step(s, m=m, chunks=chunks):
        for c in range(chunks):
            c_lo, c_hi = c * m, (n if c == chunks - 1 else (c + 1) * m)
            run(s, c_lo, c_hi, None)
            done = torch.cuda.Event(); done.record(cur)
            side.wait_event(done)
            ens.push(lo + c_lo, lo + c_hi, stream=side)
        cur.wait_stream(side)
        dist.all_reduce(flag)
    say(f"DMA push, {chunks} chunk(s): {timed(step):.2f} ms")
# correctness of the DMA path: every other rank's shard as it arrived here against a local recomputation
bad = []
for other in range(world):
    if other == rank:
        continue
    olo, ohi = sharding.shard_range(n_total, other, world)
    odraws = torch.from_numpy(bench.synthetic_draws(n_total, 1718)[:, olo:ohi].copy()).to(dev)
    want, _ = engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), ohi - olo, 0, 162, 0.5, var=(bench.VAR_INDEX, odraws), tdpf=(tab, idx),
                              seed=103, traj_id0=olo, out_dtype=torch.int32, dev=dev)
    got = ens.full[:, :, olo:ohi]
    diff = (got != want)
    bad.append((other, int(diff.sum()), int(diff.any(dim=0).any(dim=0).sum()), int(diff.any(dim=2).sum())))
    del want, diff
print(f"rank {rank}: (source rank, mismatching elements, columns, (row, time) cells) = {bad}", flush=True)
ens.close()
dist.destroy_process_group()
