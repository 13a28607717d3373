"""world_size-2 (and 4) gloo tests of the multi-GPU host logic on CPU: column -> row all-to-all, Merkle-cap all-gather
and folding, MSM chunk combination.  The compute backend here is the CPU oracle (test infrastructure); on the GPU
box the same functions run with GpuBackend over NCCL (bench.py)."""
import os, sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def lde(self, cols, w, nbits, nbits_ext):
        from oracle import gl
        cm = cols.numpy().view(np.uint64).reshape(w, 1 << nbits)
        ext = gl.lde(np.ascontiguousarray(cm.T), w, nbits, nbits_ext).reshape(1 << nbits_ext, w)
        return torch.from_numpy(np.ascontiguousarray(ext.T).reshape(-1).view(np.int64))

    def merkelize(self, cols, width, height):
        from oracle import gl
        cm = cols.numpy().view(np.uint64).reshape(width, height)
        return torch.from_numpy(gl.merkelize(np.ascontiguousarray(cm.T), width, height).reshape(-1).view(np.int64))

    def hash2(self, l, r):
        from oracle import gl
        return gl.poseidon(list(l) + list(r), [0, 0, 0, 0])[:4]

    def msm(self, bases, scalars, n):
        from oracle import bn254 as bn
        aff = bn.msm_c(bases.numpy().view(np.uint64).reshape(n, 8), scalars.numpy().view(np.uint64).reshape(n, 4))
        out = np.zeros(12, dtype=np.uint64)
        if aff.any():
            out[:8] = aff; out[8:] = bn._limbs(bn.MONT_R)
        else:
            out[4:8] = bn._limbs(bn.MONT_R)
        return out

    def g1_add(self, a, b):
        from oracle import bn254 as bn
        aff = lambda j: j[:8] if j[8:].any() else np.zeros(8, dtype=np.uint64)
        s = np.zeros(8, dtype=np.uint64)
        import ctypes
        x = np.ascontiguousarray(aff(a)); y = np.ascontiguousarray(aff(b))
        bn.lib().bn_add_affine(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p))
        out = np.zeros(12, dtype=np.uint64)
        if s.any():
            out[:8] = s; out[8:] = bn._limbs(bn.MONT_R)
        else:
            out[4:8] = bn._limbs(bn.MONT_R)
        return out


    # table mode and the device-side combine of the gathered partial sums, restated for the CPU backend
    def msm_table(self, bases_chunk, n):
        return (bases_chunk, n)

    def msm_table_run(self, table, scalars_chunk):
        return self.msm(table[0], scalars_chunk, table[1])

    def points_sum(self, gathered, count):
        pts = gathered.numpy().view(np.uint64).reshape(count, 12)
        acc = pts[0].copy()
        for k in range(1, count):
            acc = self.g1_add(acc, pts[k])
        return acc


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from eigen_zkvm_b200 import sharded
        from oracle import gl, bn254 as bn
        be = OracleBackend()
        W, nbits, nbits_ext = 8, 6, 9
        rng = np.random.default_rng(3)
        full = rng.integers(0, 2**62, size=(W, 1 << nbits), dtype=np.uint64)          # column-major [col][row]
        lo, hi = sharded.column_shard(W, world, rank)
        local = torch.from_numpy(np.ascontiguousarray(full[lo:hi]).reshape(-1).view(np.int64))
        root, nodes, shard = sharded.lde_merkle_sharded(local, W, nbits, nbits_ext, be)
        # single-process truth
        ext = gl.lde(np.ascontiguousarray(full.T), W, nbits, nbits_ext)
        truth = [int(x) for x in gl.merkelize(ext, W, 1 << nbits_ext)[-1]]
        ok1 = root == truth
        rows = (1 << nbits_ext) // world
        e2 = ext.reshape(1 << nbits_ext, W)
        ok2 = (shard.numpy().view(np.uint64).reshape(W, rows) == e2[rank * rows:(rank + 1) * rows].T).all()
        # MSM chunks
        n = 64
        import random
        rnd = random.Random(1)
        pts = [bn.mul(rnd.randrange(1, bn.R), bn.G1) for _ in range(n)]; sc = [rnd.randrange(bn.R) for _ in range(n)]
        B = torch.from_numpy(bn.pack_points(pts).reshape(-1).view(np.int64)); S = torch.from_numpy(bn.pack_scalars(sc).reshape(-1).view(np.int64))
        res = sharded.msm_sharded(B, S, n, be)
        truth_msm = bn.msm_c(bn.pack_points(pts), bn.pack_scalars(sc))
        ok3 = (res[:8] == truth_msm).all()
        lo_m, n_m = sharded.msm_chunk(n, world, rank)
        tab = be.msm_table(B[lo_m * 8:(lo_m + n_m) * 8], n_m)                       # per-rank resident table of its chunk
        res2 = sharded.msm_sharded(B, S, n, be, table=tab)
        ok3 = ok3 and (res2[:8] == truth_msm).all()
        q.put((rank, bool(ok1), bool(ok2), bool(ok3)))
    except Exception as e:      # pragma: no cover
        import traceback
        q.put((rank, "ERR", traceback.format_exc(), str(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_paths_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world + (os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    for r in res:
        assert r[1:] == (True, True, True), r


def test_fold_roots_matches_tree():
    from eigen_zkvm_b200 import sharded
    from oracle import gl
    rng = np.random.default_rng(5)
    leaves = rng.integers(0, 2**62, size=(8, 5), dtype=np.uint64)
    nodes = gl.merkelize(leaves, 5, 8)
    subs = [[int(x) for x in nodes[i]] for i in range(8)]           # level 0 digests as "sub-roots"
    assert sharded.fold_roots(subs, OracleBackend().hash2) == [int(x) for x in nodes[-1]]
