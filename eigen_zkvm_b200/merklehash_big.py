"""Host-side mirror of the reference's BN128 / BLS12-381 commitment back-ends (`MerkleTreeBN128`, `MerkleTreeBLS12381`,
`LinearHashBN128`, `Poseidon`; starky/src/merklehash_bn128.rs, linearhash_bn128.rs, poseidon_bn128_opt.rs and twins)
over the C-ABI.  Field elements and digests are python ints in [0, r) at this level (4 x u64 canonical on the wire)."""
import ctypes
import numpy as np
from . import _lib

FIELD_IDS = {"BN128": 0, "BLS12381": 1, "bn128": 0, "bls12381": 1}


def _fid(field):
    if field not in FIELD_IDS:
        raise ValueError("unknown verificationHashType %r" % (field,))
    return FIELD_IDS[field]


def _to_words(vals):
    return np.array([[(int(v) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64).reshape(-1, 4)


def _from_words(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(r[i]) << (64 * i) for i in range(4)) for r in a]


class Poseidon:
    """poseidon_bn128_opt.rs:80-118 / poseidon_bls12381_opt.rs:81-113"""
    def __init__(self, field):
        self.fid = _fid(field)

    def hash_ex(self, inp, init_state, out):
        if len(inp) == 0 or len(inp) > 16:
            raise ValueError("Wrong inputs length %d > 16" % len(inp))
        i = _to_words(inp); s = _to_words([init_state]); o = np.zeros((len(inp) + 1, 4), dtype=np.uint64)
        _lib.check(_lib.lib().b200_big_poseidon(self.fid, i.ctypes.data_as(ctypes.c_void_p), len(inp), s.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p)))
        return _from_words(o)[:out]

    def hash(self, inp, init_state=0):
        if len(inp) == 0 or len(inp) > 16:
            raise ValueError("Wrong inputs length %d > 16" % len(inp))
        i = _to_words(inp); s = _to_words([init_state]); o = np.zeros(4, dtype=np.uint64)
        _lib.check(_lib.lib().b200_big_hash(self.fid, i.ctypes.data_as(ctypes.c_void_p), len(inp), s.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p)))
        return _from_words(o)[0]


class LinearHash:
    def __init__(self, field):
        self.fid = _fid(field)

    def hash_element_array(self, rows, width):
        """digests of n rows of `width` GL elements (row-major u64 buffer)"""
        r = np.ascontiguousarray(rows, dtype=np.uint64).reshape(-1, width) if width else np.zeros((len(rows), 0), dtype=np.uint64)
        o = np.zeros((r.shape[0], 4), dtype=np.uint64)
        _lib.check(_lib.lib().b200_big_linearhash(self.fid, r.ctypes.data_as(ctypes.c_void_p), width, r.shape[0], o.ctypes.data_as(ctypes.c_void_p)))
        return _from_words(o)


class MerkleTree:
    """`MerkleTree` trait subset (starky/src/traits.rs:24-55): new, merkelize(buff, width, height), root, get_group_proof."""
    def __init__(self, field):
        self.fid = _fid(field); self.nodes = []; self.elements = None; self.width = 0; self.height = 0

    def merkelize(self, buff, width, height):
        b = np.ascontiguousarray(buff, dtype=np.uint64)
        if width and b.size != width * height:
            raise ValueError("buffer size does not match width * height")
        nn = _lib.lib().b200_big_merkle_n_nodes(height)
        o = np.zeros((nn, 4), dtype=np.uint64)
        _lib.check(_lib.lib().b200_big_merkelize(self.fid, b.ctypes.data_as(ctypes.c_void_p), width, height, o.ctypes.data_as(ctypes.c_void_p)))
        self.nodes = _from_words(o); self.elements = b.reshape(height, width) if width else b; self.width = width; self.height = height

    def root(self):
        return self.nodes[-1]

    def get_group_proof(self, idx):
        """merklehash_bn128.rs:89-106,226-243: (row values, [16 siblings per level])"""
        if idx >= self.height:
            raise IndexError("access invalid node")
        v = [int(x) for x in self.elements[idx]]
        mp = []; n = self.height; offset = 0
        while n > 1:
            si = idx & ~0xF
            mp.append(self.nodes[offset + si: offset + si + 16])
            nn = (n - 1) // 16 + 1
            offset += nn * 16; idx >>= 4; n = nn
        return v, mp
