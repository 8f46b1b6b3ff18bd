"""Worker of the fused-gather test: `python dist_peer_worker.py <rank> <world> <port> <backend> <out.npz>`.
Every rank simulates its shard of ONE influenza ensemble on cuda:(rank % device_count); the tau-leap kernel stores each output
row into the rank's own copy of the ensemble and into the other ranks' copies (`sharding.PeerEnsemble`, CUDA IPC mappings).
Rank 0 writes its full copy for comparison with the single-process run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, backend, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend, rank=rank, world_size=world)
    import cases
    from pysodm_b200 import engine, sharding
    g = np.load(os.path.join(ROOT, "tests", "golden", "jump_single.npz"))
    dev = engine.device(rank % torch.cuda.device_count())
    n_total = 3001                                              # odd: shards of different length, unaligned column offsets
    ens = sharding.PeerEnsemble(28, cases.FLU_N_OUT, n_total, torch.int32, dev)
    assert len(ens.peer_columns) == world - 1
    ens.full.fill_(-1)
    ens.fence()
    n = ens.hi - ens.lo
    _, st = engine.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5,
                            tdpf=(g["flu_table"], g["flu_step_index"]), seed=99, traj_id0=ens.lo, out_dtype=torch.int32,
                            out=ens.local, out_peers=ens.peer_columns, dev=dev)
    ens.fence()
    assert int(st.abs().sum()) == 0
    full = ens.full.cpu().numpy()
    sums = [None] * world
    dist.all_gather_object(sums, int(full.astype(np.int64).sum()))
    assert len(set(sums)) == 1, sums                            # every rank holds the same ensemble
    if rank == 0:
        np.savez(path, full=full)
    ens.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
