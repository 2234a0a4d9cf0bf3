"""Worker of tests/test_gpu_nccl.py: one rank of a row-tiled solve over NCCL (one GPU per rank)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch.distributed as dist  # noqa: E402  (plumbing only: hands out the NCCL unique id)

import oracle  # noqa: E402
from pyflwdir_b200 import tiled  # noqa: E402
from pyflwdir_b200.pyflwdir import _get_idxs_dtype  # noqa: E402

out_dir = sys.argv[1]
dist.init_process_group(backend="gloo")
rank, world = dist.get_rank(), dist.get_world_size()
z = oracle.synth_elevation(640, 384, seed=17)
d8 = oracle.synth_d8(z, sea_level=float(np.quantile(z, 0.05)))
blocks = tiled.split_rows(d8.shape[0], world)
assert len(blocks) == world
solver = tiled.RowBlockSolver(int(os.environ.get("LOCAL_RANK", "0")))
uid = [tiled.RowBlockSolver.unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
solver.comm_init(rank, world, uid[0])
r0, r1 = blocks[rank]
blk, ht, hb = tiled.block_with_halo(d8, r0, r1)
ids, rk, upa, bas, npits = solver.flow_all(blk, ht, hb, r0, _get_idxs_dtype(d8.size))
np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=ids, rk=rk, upa=upa, bas=bas, npits=npits)
dist.barrier()
if rank == 0:
    parts = [np.load(os.path.join(out_dir, f"rank{g}.npz")) for g in range(world)]
    ids_o, pits_o, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
    seq = oracle.core.idxs_seq(ids_o, pits_o)
    upa_o = oracle.streams.accuflux(ids_o, seq, np.ones(d8.size, np.int32), -9999)
    upa_o[ids_o == -1] = -9999
    assert int(parts[0]["npits"]) == pits_o.size
    assert np.array_equal(np.concatenate([p["ids"] for p in parts]), ids_o)
    assert np.array_equal(np.concatenate([p["rk"] for p in parts]).ravel(), oracle.core.rank(ids_o)[0])
    assert np.array_equal(np.concatenate([p["upa"] for p in parts]).ravel(), upa_o)
    assert np.array_equal(np.concatenate([p["bas"] for p in parts]).ravel(), oracle.basins.basins(ids_o, pits_o, seq))
    print("NCCL_TILED_OK", world, flush=True)
dist.barrier()
# the order-sensitive outputs across the row blocks: halo rounds over ncclSend / ncclRecv between row neighbours
area = (np.abs(z) + np.float32(0.25)).astype(np.float32)
so, rounds = solver.sweep("strahler")
acc, _ = solver.sweep("accuflux", data=area[r0:r1], nodata=-9999.0)
upa_all = None
np.savez(os.path.join(out_dir, f"sweep{rank}.npz"), so=so, acc=acc)
dist.barrier()
parts = [np.load(os.path.join(out_dir, f"sweep{g}.npz")) for g in range(world)]
ids_o, pits_o, _ = oracle.core_d8.from_array(d8, dtype=np.int32)
seq_o = oracle.core.idxs_seq(ids_o, pits_o)
upa_o = oracle.streams.accuflux(ids_o, seq_o, np.ones(d8.size, np.int32), -9999).reshape(d8.shape)
drain = upa_o > 60
hand, _ = solver.sweep("hand", data=z[r0:r1], drain=drain[r0:r1])
np.save(os.path.join(out_dir, f"hand{rank}.npy"), hand)
dist.barrier()
if rank == 0:
    assert np.array_equal(np.concatenate([p["so"] for p in parts]).ravel(), oracle.streams.strahler_order(ids_o, seq_o))
    assert np.array_equal(np.concatenate([p["acc"] for p in parts]).ravel(), oracle.streams.accuflux(ids_o, seq_o, area.ravel(), -9999.0))
    hands = np.concatenate([np.load(os.path.join(out_dir, f"hand{g}.npy")) for g in range(world)])
    assert np.array_equal(hands.ravel(), oracle.dem.height_above_nearest_drain(ids_o, seq_o, drain.ravel(), z.ravel()))
    assert rounds >= 2
    print("NCCL_SWEEPS_OK", world, rounds, flush=True)
dist.barrier()
# error agreement: ONE rank's block holds an illegal code -> every rank returns an error instead of hanging in a collective
bad = d8.copy()
rb0, rb1 = blocks[world - 1]
bad[(rb0 + rb1) // 2, 7] = 3
blk, ht, hb = tiled.block_with_halo(bad, r0, r1)
try:
    solver.flow_all(blk, ht, hb, r0, _get_idxs_dtype(d8.size))
    raised = None
except ValueError as err:
    raised = str(err)
oks = [None] * world
dist.all_gather_object(oks, raised)
if rank == 0:
    assert all(o is not None and "D8 code set" in o for o in oks), oks
    print("NCCL_ERROR_AGREEMENT_OK", world, flush=True)
# ... and the communicator is still usable afterwards
ids2, rk2, upa2, bas2, _ = solver.flow_all(*tiled.block_with_halo(d8, r0, r1), r0, _get_idxs_dtype(d8.size))
assert np.array_equal(ids2, ids) and np.array_equal(rk2, rk) and np.array_equal(upa2, upa) and np.array_equal(bas2, bas)
dist.barrier()
solver.close()
dist.destroy_process_group()
