"""ctypes front-end of the CPU oracle (oracle/gl_oracle.c) -- TEST INFRASTRUCTURE, not product code.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Everything is canonical u64 numpy arrays, row-major [row][col] like the reference buffers
(starky/src/polsarray.rs:219-227).
"""
import ctypes, os, subprocess
import numpy as np

P = 0xFFFFFFFF00000001
SHIFT = 49
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    """Compile the C restatement (gcc + OpenMP). Idempotent."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libgl_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        u64, sz, vp = ctypes.c_uint64, ctypes.c_size_t, ctypes.c_void_p
        for name in ("ora_gl_add", "ora_gl_sub", "ora_gl_mul", "ora_gl_mul_slow", "ora_gl_pow"):
            getattr(L, name).restype = u64
            getattr(L, name).argtypes = [u64, u64]
        L.ora_gl_inv.restype = u64
        L.ora_gl_inv.argtypes = [u64]
        L.ora_root.restype = u64
        L.ora_root.argtypes = [ctypes.c_uint]
        L.ora_root_inv.restype = u64
        L.ora_root_inv.argtypes = [ctypes.c_uint]
        L.ora_merkle_n_nodes.restype = sz
        L.ora_merkle_n_nodes.argtypes = [sz]
        L.ora_merkle_proof.restype = sz
        L.ora_merkle_proof.argtypes = [vp, sz, sz, vp]
        L.ora_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _u(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


# ---- scalars (python ints) -------------------------------------------------------------------
def add(a, b): return (a + b) % P
def sub(a, b): return (a - b) % P
def mul(a, b): return (a * b) % P
def inv(a): return pow(a, P - 2, P)
def root(k): return int(lib().ora_root(k))
def root_inv(k): return int(lib().ora_root_inv(k))


# ---- F3G on python tuples: GL[x]/(x^3 - x - 1)  (starky/src/f3g.rs:407-449) -------------------
def f3(a, b=0, c=0): return (a % P, b % P, c % P)
def f3_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P, (a[2] + b[2]) % P)
def f3_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P, (a[2] - b[2]) % P)
def f3_neg(a): return ((-a[0]) % P, (-a[1]) % P, (-a[2]) % P)
def f3_muls(a, s): return (a[0] * s % P, a[1] * s % P, a[2] * s % P)


def f3_mul(a, b):
    A = (a[0] + a[1]) * (b[0] + b[1]); B = (a[0] + a[2]) * (b[0] + b[2]); C = (a[1] + a[2]) * (b[1] + b[2])
    D = a[0] * b[0]; E = a[1] * b[1]; F = a[2] * b[2]; G = D - E
    return ((C + G - F) % P, (A + C - E - E - D) % P, (B - G) % P)


def f3_inv(x):
    out = np.zeros(3, dtype=np.uint64)
    lib().ora_f3_inv(_p(_u(list(x))), _p(out))
    return tuple(int(v) for v in out)


def f3_pow(a, e):
    r = (1, 0, 0)
    while e:
        if e & 1: r = f3_mul(r, a)
        a = f3_mul(a, a); e >>= 1
    return r


def f3_div(a, b): return f3_mul(a, f3_inv(b))


# ---- hashing ----------------------------------------------------------------------------------
def poseidon(in8, cap4):
    """Poseidon t=12 permutation output, all 12 lanes (poseidon_opt.rs:80-200)."""
    out = np.zeros(12, dtype=np.uint64)
    lib().ora_poseidon(_p(_u(in8)), _p(_u(cap4)), _p(out))
    return [int(v) for v in out]


def linearhash(vals):
    v = _u(vals)
    out = np.zeros(4, dtype=np.uint64)
    lib().ora_linearhash(_p(v), ctypes.c_size_t(v.size), _p(out))
    return [int(x) for x in out]


def merkle_n_nodes(h): return int(lib().ora_merkle_n_nodes(h))


def merkelize(leaves, width, height):
    """nodes array (n_nodes x 4), layout of MerkleTreeGL.nodes (merklehash.rs:293-346)."""
    leaves = _u(leaves).reshape(-1)
    assert leaves.size == width * height
    nodes = np.zeros((merkle_n_nodes(height), 4), dtype=np.uint64)
    lib().ora_merkelize(_p(leaves), ctypes.c_size_t(width), ctypes.c_size_t(height), _p(nodes))
    return nodes


def merkle_proof(nodes, height, idx):
    buf = np.zeros((64, 4), dtype=np.uint64)
    d = lib().ora_merkle_proof(_p(nodes), height, idx, _p(buf))
    return buf[:d].copy()


def merkle_root_from_proof(vals, sibs, idx):
    v = _u(vals); s = _u(sibs)
    out = np.zeros(4, dtype=np.uint64)
    lib().ora_merkle_root_from_proof(_p(v), ctypes.c_size_t(v.size), _p(s), ctypes.c_size_t(s.shape[0] if s.ndim == 2 else s.size // 4),
                                     ctypes.c_size_t(idx), _p(out))
    return [int(x) for x in out]


# ---- polynomials ------------------------------------------------------------------------------
def ntt(a, w, bits):
    a = _u(a); out = np.empty_like(a)
    lib().ora_ntt(_p(a), _p(out), ctypes.c_size_t(w), ctypes.c_uint(bits)); return out


def intt(a, w, bits):
    a = _u(a); out = np.empty_like(a)
    lib().ora_intt(_p(a), _p(out), ctypes.c_size_t(w), ctypes.c_uint(bits)); return out


def lde(a, w, bits, bits_ext):
    a = _u(a).reshape(-1)
    out = np.zeros((1 << bits_ext) * w, dtype=np.uint64)
    if w:
        lib().ora_lde(_p(a), _p(out), ctypes.c_size_t(w), ctypes.c_uint(bits), ctypes.c_uint(bits_ext))
    return out


def f3_batch_inverse(a):
    a = _u(a).reshape(-1); out = np.empty_like(a)
    lib().ora_f3_batch_inverse(_p(a), _p(out), ctypes.c_size_t(a.size // 3)); return out
