"""Multi-GPU sharding of the two paths that shard, one process per GPU over torch.distributed.

* Wide traces (BASELINE configs[2]): NTT/LDE shards by independent columns with no communication; leaf hashing needs
  whole rows, so one all-to-all turns column shards into row shards; every rank builds the Merkle subtree over its
  rows and the sub-roots (the "Merkle cap", 32 B each) are all-gathered and folded by every rank.
* MSM (configs[3]): independent (point, scalar) chunks; the 96-byte partial sums are all-gathered and added locally
  (point addition is not an NCCL reduction op).

The compute is behind a small backend interface so that the host logic (layout, collectives, folding) is also
exercised on CPU with gloo in tests/ (with the oracle as the compute backend); the product backend is `GpuBackend`,
which calls the C-ABI on device pointers.
"""
import ctypes
import numpy as np
import torch
import torch.distributed as dist


def column_shard(width, world, rank):
    if width % world:
        raise ValueError("width must be divisible by the number of ranks")
    per = width // world
    return rank * per, (rank + 1) * per


class GpuBackend:
    """C-ABI calls on CUDA tensors (int64 storage of canonical u64, column-major)."""
    def __init__(self):
        from . import _lib
        self._lib = _lib; self.L = _lib.lib()

    def lde(self, cols, w, nbits, nbits_ext):
        out = torch.empty(w << nbits_ext, dtype=torch.int64, device=cols.device)
        self._lib.check(self.L.b200_gl_lde_dev(ctypes.c_void_p(cols.data_ptr()), ctypes.c_void_p(out.data_ptr()), w, nbits, nbits_ext))
        return out

    def merkelize(self, cols, width, height):
        nodes = torch.empty(self.L.b200_gl_merkle_n_nodes(height) * 4, dtype=torch.int64, device=cols.device)
        self._lib.check(self.L.b200_gl_merkelize_dev(ctypes.c_void_p(cols.data_ptr()), width, height, ctypes.c_void_p(nodes.data_ptr())))
        return nodes

    def hash2(self, left4, right4):
        i8 = np.array(list(left4) + list(right4), dtype=np.uint64); c4 = np.zeros(4, dtype=np.uint64); o = np.zeros(12, dtype=np.uint64)
        self._lib.check(self.L.b200_gl_poseidon(i8.ctypes.data_as(ctypes.c_void_p), c4.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p)))
        return [int(x) for x in o[:4]]

    def msm(self, bases, scalars, n):
        from . import groth16
        return groth16.multiexp_dev(bases.data_ptr(), scalars.data_ptr(), n)

    def g1_add(self, a, b):
        from . import groth16
        return groth16.g1_add(a, b)

    # per-circuit table of this rank's chunk of the bases (resident; built once)
    def msm_table(self, bases_chunk, n):
        from . import groth16
        t = groth16.MsmTable(device_ptr=bases_chunk.data_ptr(), n=n)
        if dist.is_initialized() and dist.get_world_size() > 1:
            t.set_partial_output(True)      # a rank's partial sum stays un-normalised: only the combined point is inverted (points_sum)
        return t

    def msm_table_run(self, table, scalars_chunk):
        return table.run_dev(scalars_chunk.data_ptr())

    def points_sum(self, gathered, count):
        from . import groth16
        return groth16.points_sum_dev(gathered.data_ptr(), count)


def fold_roots(roots, hash2):
    """Binary Merkle levels over the per-rank sub-roots (merklehash.rs:79-134 applied to the top log2(world) levels)."""
    lvl = [list(r) for r in roots]
    while len(lvl) > 1:
        if len(lvl) & 1:
            lvl.append([0, 0, 0, 0])
        lvl = [hash2(lvl[2 * i], lvl[2 * i + 1]) for i in range(len(lvl) // 2)]
    return lvl[0]


def lde_merkle_sharded(local_cols, width, nbits, nbits_ext, backend, group=None):
    """local_cols: this rank's columns [width/world][2^nbits] (column-major, int64 tensor).
    Returns (root, subtree_nodes, row_shard) where row_shard is [width][2^nbits_ext/world] column-major."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = column_shard(width, world, rank)
    w_local = hi - lo
    n_ext = 1 << nbits_ext
    if n_ext % world or (world & (world - 1)):
        raise ValueError("world size must be a power of two dividing the extended height")
    rows = n_ext // world
    ext = backend.lde(local_cols, w_local, nbits, nbits_ext)                       # [w_local][n_ext], no communication
    if world > 1:
        send = ext.view(w_local, world, rows).permute(1, 0, 2).contiguous()        # [dest][w_local][rows]
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)          # column shards -> row shards
        # The library launches on its own stream handle; torch only orders ITS current stream behind the NCCL stream
        # (a race seen on 2 x B200 in round 1).  Join on an EVENT recorded behind the collective on torch's stream -- the host waits
        # for the exchange only, not for everything else in flight on the device.
        if recv.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(recv.device))
            ev.synchronize()
        row_shard = recv.view(-1)                                                  # [src][w_local][rows] == [width][rows]
    else:
        row_shard = ext
    nodes = backend.merkelize(row_shard, width, rows)                              # subtree over this rank's rows
    sub = nodes[-4:].cpu().numpy().view(np.uint64) if nodes.dtype == torch.int64 else nodes[-4:]
    sub = [int(x) for x in sub]
    if world > 1:
        t = torch.tensor(np.array(sub, dtype=np.uint64).view(np.int64), device=nodes.device)
        allr = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allr, t, group=group)                                      # the Merkle cap: world x 32 B
        roots = [[int(x) for x in a.cpu().numpy().view(np.uint64)] for a in allr]
    else:
        roots = [sub]
    return fold_roots(roots, backend.hash2), nodes, row_shard


def msm_chunk(n_total, world, rank):
    per = n_total // world
    lo = rank * per
    return lo, (per if rank < world - 1 else n_total - lo)


def msm_sharded(bases, scalars, n_total, backend, group=None, table=None):
    """bases / scalars: the FULL device arrays (each rank reads its own chunk); returns the combined (X, Y, Z).
    table: this rank's resident table (backend.msm_table of its chunk of the bases): the window width is then chosen from the
    per-rank chunk and there is one bucket set per rank, so accumulation AND bucket reduction shrink with the rank count.
    The partial sums (96 B each) are all-gathered into one device buffer and added by one kernel on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, n_mine = msm_chunk(n_total, world, rank)
    if table is not None:
        part = backend.msm_table_run(table, scalars[lo * 4:(lo + n_mine) * 4])
    else:
        part = backend.msm(bases[lo * 8:(lo + n_mine) * 8], scalars[lo * 4:(lo + n_mine) * 4], n_mine)
    if world == 1:
        return part
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.int64)).to(bases.device)
    gathered = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(gathered, t, group=group)
    if gathered.is_cuda:
        torch.cuda.current_stream(gathered.device).synchronize()      # the library runs on its own stream: join behind the collective (one 768-byte buffer)
    return backend.points_sum(gathered, world)
