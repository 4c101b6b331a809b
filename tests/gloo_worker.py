"""Worker of tests/test_dist_gloo.py: the host transport of the multi-GPU wiring on CPU (gloo, no GPU)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from petibm_b200.dist import Comm  # noqa: E402
from petibm_b200.mesh import slab_range  # noqa: E402


def main():
    comm = Comm.from_env(backend="gloo")
    assert comm.nranks == int(os.environ["WORLD_SIZE"]) and comm.rank == int(os.environ["RANK"])
    # 64-byte "IPC handles" travel rank-major
    mine = bytes([comm.rank]) * 64
    got = comm.allgather_bytes(mine)
    assert got == [bytes([r]) * 64 for r in range(comm.nranks)]
    uid = comm.broadcast_bytes(b"\x07" * 128 if comm.rank == 0 else None, 0)
    assert uid == b"\x07" * 128
    assert comm.allreduce_max(float(comm.rank)) == float(comm.nranks - 1)
    # natural-ordering vector <-> DMDA z-slabs
    n = (5, 4, 7)
    full = np.arange(np.prod(n), dtype=np.float64)
    loc = comm.local_block(full, n)
    lo, hi = slab_range(n[2], comm.rank, comm.nranks)
    assert loc.size == n[0] * n[1] * (hi - lo) and loc[0] == lo * n[0] * n[1]
    back = comm.gather_blocks(loc)
    assert np.array_equal(back, full)
    # 2-D grids are cut into y-slabs
    full2 = np.arange(6 * 9, dtype=np.float64)
    loc2 = comm.local_block(full2, (6, 9))
    lo2, hi2 = slab_range(9, comm.rank, comm.nranks)
    assert loc2.size == 6 * (hi2 - lo2) and loc2[0] == lo2 * 6
    assert np.array_equal(comm.gather_blocks(loc2), full2)
    # DMDA boxes <-> solver slabs through ONE all-to-all per direction (tests/test_repartition.py checks the plan
    # itself; this is the transport): 2 x 1 x 1 or 2 x 2 x 1 process grid, uneven cuts
    from petibm_b200.dist import Repart

    n3 = (7, 6, 9)
    procs = {2: (2, 1, 1), 4: (2, 2, 1)}.get(comm.nranks, (1, 1, comm.nranks))
    plan = Repart(3, n3, procs, comm.rank)
    field = 100.0 + np.arange(np.prod(n3), dtype=np.float64)
    box = field[plan.box_rows()]
    slab = plan.box_to_slab(box)
    lo3, hi3 = plan.slab
    assert np.array_equal(slab, field[lo3 * 42: hi3 * 42])
    assert np.array_equal(plan.slab_to_box(slab), box)
    # a row-distributed matrix (uneven blocks, global column indices) gathered on every rank: the replicated solve
    import scipy.sparse as sp

    rng = np.random.default_rng(5)
    ntot = 23
    M = (sp.random(ntot, ntot, density=0.3, random_state=7, format="csr") + sp.identity(ntot)).tocsr()
    M.sort_indices()
    cuts = np.linspace(0, ntot, comm.nranks + 1).astype(int)
    cuts[1:-1] += (np.arange(1, comm.nranks) % 2)            # uneven
    mine_rows = M[cuts[comm.rank]: cuts[comm.rank + 1]]
    vec = rng.standard_normal(ntot)
    ip, ix, dv, offs, hc, nv = comm.gather_matrix(mine_rows.shape[0], mine_rows.indptr, mine_rows.indices, mine_rows.data, False,
                                                  vec[cuts[comm.rank]: cuts[comm.rank + 1]])
    assert np.array_equal(offs, cuts) and not hc
    assert np.array_equal(ip, M.indptr) and np.array_equal(ix, M.indices) and np.array_equal(dv, M.data)
    assert nv.shape == (1, ntot) and np.array_equal(nv[0], vec)
    # the per-solve transport of the replicated solves: uneven float64 parts, rank order
    assert np.array_equal(comm.allgather_f64(vec[cuts[comm.rank]: cuts[comm.rank + 1]]), vec)
    comm.barrier()
    import torch.distributed as dist

    dist.destroy_process_group()
    print(f"GLOO_WORKER_OK rank {comm.rank}")


if __name__ == "__main__":
    main()
