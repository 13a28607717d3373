"""Host-side mirror of the reference's starky interface over the C-ABI: same names, argument meaning and
error behaviour as the Rust functions they stand for, so the parity tests read like the reference's own tests.

  fft / ifft / interpolate      starky/src/fft_p.rs:242-261
  Poseidon.hash                 starky/src/poseidon_opt.rs:76-78
  LinearHash.hash               starky/src/linearhash.rs:79-110
  MerkleTreeGL                  starky/src/merklehash.rs:36-45,255-467 (trait: traits.rs:24-55)
  StarkSetup.new                starky/src/stark_setup.rs:27-66
  StarkProof.stark_gen          starky/src/stark_gen.rs:193-202
  stark_prove                   starky/src/prove.rs:30-91 (GL hash type)

Buffers are numpy uint64 arrays of canonical field values, row-major [row][col] like `PolsArray.write_buff`.
"""
import ctypes, json
import numpy as np
from . import _lib
from . import starkinfo as _si

P = 0xFFFFFFFF00000001


def _arr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
    return a


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---- fft_p.rs ---------------------------------------------------------------------------------------------
def fft(buffsrc, n_pols, nbits):
    src = _arr(buffsrc); assert src.size == n_pols << nbits
    dst = np.empty_like(src)
    _lib.check(_lib.lib().b200_gl_ntt(_ptr(src), _ptr(dst), n_pols, nbits))
    return dst


def ifft(buffsrc, n_pols, nbits):
    src = _arr(buffsrc); assert src.size == n_pols << nbits
    dst = np.empty_like(src)
    _lib.check(_lib.lib().b200_gl_intt(_ptr(src), _ptr(dst), n_pols, nbits))
    return dst


def interpolate(buffsrc, n_pols, nbits, nbitsext):
    src = _arr(buffsrc); assert src.size == n_pols << nbits
    dst = np.zeros(n_pols << nbitsext, dtype=np.uint64)
    if src.size == 0:
        return dst          # fft_p.rs:262-264
    _lib.check(_lib.lib().b200_gl_lde(_ptr(src), _ptr(dst), n_pols, nbits, nbitsext))
    return dst


# ---- hashing ----------------------------------------------------------------------------------------------
class Poseidon:
    def hash(self, inp, init_state, out=4):
        if len(inp) != 8: raise ValueError("Wrong inputs length %d != 8" % len(inp))
        if len(init_state) != 4: raise ValueError("Capacity inputs length %d != 4" % len(init_state))
        a = _arr(inp); c = _arr(init_state); o = np.zeros(12, dtype=np.uint64)
        _lib.check(_lib.lib().b200_gl_poseidon(_ptr(a), _ptr(c), _ptr(o)))
        return [int(x) for x in o[:out]]


class LinearHash:
    def hash(self, flatvals, batch_size=0):
        assert batch_size == 0
        v = _arr(flatvals); o = np.zeros(4, dtype=np.uint64)
        _lib.check(_lib.lib().b200_gl_linearhash(_ptr(v), v.size, 1, _ptr(o)))
        return [int(x) for x in o]

    def hash_rows(self, rows, width):
        v = _arr(rows); n = v.size // width if width else 0
        o = np.zeros((n, 4), dtype=np.uint64)
        _lib.check(_lib.lib().b200_gl_linearhash(_ptr(v), width, n, _ptr(o)))
        return o


def get_n_nodes(n):
    return int(_lib.lib().b200_gl_merkle_n_nodes(n))


class MerkleTreeGL:
    def __init__(self):
        self.elements = np.zeros(0, dtype=np.uint64); self.width = 0; self.height = 0
        self.nodes = np.zeros((0, 4), dtype=np.uint64)

    def merkelize(self, buff, width, height):
        self.elements = _arr(buff); self.width, self.height = width, height
        assert self.elements.size == width * height
        self.nodes = np.zeros((get_n_nodes(height), 4), dtype=np.uint64)
        _lib.check(_lib.lib().b200_gl_merkelize(_ptr(self.elements), width, height, _ptr(self.nodes)))

    def root(self):
        return [int(x) for x in self.nodes[-1]]

    def get_element(self, idx, sub_idx):
        return int(self.elements[self.width * idx + sub_idx])

    def get_group_proof(self, idx):
        if idx >= self.height:
            raise IndexError("MerkleTreeError: access invalid node")
        v = [self.get_element(idx, i) for i in range(self.width)]
        mp = []; n = self.height; off = 0
        while n > 1:
            mp.append([int(x) for x in self.nodes[off + (idx ^ 1)]])
            nn = (n - 1) // 2 + 1
            off += nn * 2; idx >>= 1; n = nn
        return v, mp


# ---- StarkSetup / StarkProof --------------------------------------------------------------------------------
class StarkSetup:
    """StarkSetup::new: const-pol LDE + const tree on the device, PIL codegen on the host."""
    def __init__(self):
        self._h = None; self.starkinfo = None; self.program = None; self.const_root = None; self.stark_struct = None

    @classmethod
    def new(cls, const_pols, pil, stark_struct, global_l1=None):
        self = cls()
        pil = _si.load_pil(pil) if not (isinstance(pil, dict) and "cm_dims" in pil) else pil
        self.starkinfo, self.program = _si.new_starkinfo(pil, stark_struct, global_l1)
        self.stark_struct = stark_struct
        self.pil = pil
        c = _arr(const_pols)
        n = 1 << stark_struct["nBits"]
        nconst = self.starkinfo.n_constants
        if c.size != n * nconst:
            raise ValueError("const_pol.nPols != pil.nConstants")
        js = _si.setup_json(self.starkinfo, self.program, stark_struct).encode()
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().b200_setup_new(js, _ptr(c), n, nconst, ctypes.byref(h)))
        self._h = h
        r = np.zeros(4, dtype=np.uint64)
        _lib.check(_lib.lib().b200_setup_const_root(self._h, _ptr(r)))
        if stark_struct["verificationHashType"] == "GL":
            self.const_root = [int(x) for x in r]
        else:       # BN128 / BLS12381: the digest is one scalar (4 canonical u64 limbs on the wire)
            self.const_root = [sum(int(x) << (64 * i) for i, x in enumerate(r))]
        return self

    def _read_root(self):
        r = np.zeros(4, dtype=np.uint64)
        _lib.check(_lib.lib().b200_setup_const_root(self._h, _ptr(r)))
        if self.stark_struct["verificationHashType"] == "GL":
            self.const_root = [int(x) for x in r]
        else:
            self.const_root = [sum(int(x) << (64 * i) for i, x in enumerate(r))]

    def export(self, path):
        """the serialized StarkSetup (stark_setup.rs:13-19): const polynomials, their extension, the tree and the codegen output"""
        _lib.check(_lib.lib().b200_setup_export(self._h, path.encode()))

    @classmethod
    def load(cls, path, stark_struct):
        """a setup exported by `export`: nothing is recomputed (no LDE, no hashing, no codegen)"""
        self = cls()
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().b200_setup_import(path.encode(), ctypes.byref(h)))
        self._h = h; self.stark_struct = stark_struct
        shp = (ctypes.c_size_t * 4)()
        _lib.check(_lib.lib().b200_setup_shape(self._h, shp))
        if (int(shp[0]), int(shp[1])) != (stark_struct["nBits"], stark_struct["nBitsExt"]):
            raise ValueError("the file holds a setup for another stark struct")
        self.n_cm1 = int(shp[2])
        self._read_root()
        return self

    def set_self_verify(self, on=True):
        """prove.rs:124-132: make stark_gen verify every proof before returning it (raises when the verifier rejects)"""
        _lib.check(_lib.lib().b200_setup_set_self_verify(self._h, 1 if on else 0))

    def free(self):
        if self._h is not None:
            _lib.lib().b200_setup_free(self._h); self._h = None

    def __del__(self):
        try: self.free()
        except Exception: pass


class StarkProof:
    @staticmethod
    def stark_gen(cm_pols, setup, prover_addr="", device_ptr=None, n_rows=None, n_cols=None):
        """Returns the proof as the JSON string `serde_json::to_string(&starkproof)` would produce (prove.rs:153).
        cm_pols: row-major N x n_cm1 canonical u64 (host numpy), or pass device_ptr (int) of the same layout."""
        L = _lib.lib()
        out = ctypes.c_void_p(); ln = ctypes.c_size_t()
        if device_ptr is not None:
            _lib.check(L.b200_stark_gen_dev(setup._h, ctypes.c_void_p(device_ptr), n_rows, n_cols, prover_addr.encode(), ctypes.byref(out), ctypes.byref(ln)))
        else:
            cm = _arr(cm_pols)
            n = 1 << setup.stark_struct["nBits"]
            w = setup.starkinfo.n_cm1 if setup.starkinfo is not None else setup.n_cm1
            if cm.size != n * w:
                raise ValueError("cm_pols shape does not match the setup")
            _lib.check(L.b200_stark_gen(setup._h, _ptr(cm), n, w, prover_addr.encode(), ctypes.byref(out), ctypes.byref(ln)))
        return _lib.take_string(out, ln)


def stark_verify(proof_json, const_root, starkinfo, stark_struct, program, reason=None):
    """stark_verify (starky/src/stark_verify.rs:21-121): the library's host-side verifier (csrc/verify.cpp).  proof_json: the string
    stark_gen returns; const_root: StarkSetup.const_root (4 lanes for GL, [scalar] for BN128 / BLS12381).  Returns bool; `reason`
    (a list) receives the first failed check.  Goldilocks proofs verify without a GPU."""
    js = _si.setup_json(starkinfo, program, stark_struct).encode()
    if stark_struct["verificationHashType"] == "GL":
        r = np.array([int(x) for x in const_root], dtype=np.uint64)
    else:
        v = int(const_root[0]); r = np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    ok = ctypes.c_int(0); why = ctypes.c_void_p()
    pj = proof_json.encode() if isinstance(proof_json, str) else bytes(proof_json)
    _lib.check(_lib.lib().b200_stark_verify(js, _ptr(r), pj, ctypes.byref(ok), ctypes.byref(why)))
    msg = _lib.take_string(why, None)
    if reason is not None and msg:
        reason.append(msg)
    return bool(ok.value)


def stark_prove(stark_struct_file, pil_file, const_pol_file, cm_pol_file, zkin_file, prover_addr=""):
    """prove.rs:30-91 (verificationHashType GL, BN128 or BLS12381): load files -> setup -> stark_gen -> write zkin JSON."""
    ss = json.load(open(stark_struct_file))
    if ss["verificationHashType"] not in ("GL", "BN128", "BLS12381"):
        raise ValueError("Invalid hashtype %s" % ss["verificationHashType"])        # prove.rs:90
    pil = _si.load_pil(pil_file)
    const = np.fromfile(const_pol_file, dtype="<u8"); cm = np.fromfile(cm_pol_file, dtype="<u8")
    setup = StarkSetup.new(const, pil, ss)
    js = StarkProof.stark_gen(cm, setup, prover_addr)
    why = []
    if not stark_verify(js, setup.const_root, setup.starkinfo, ss, setup.program, why):      # prove.rs:124-132 asserts the same
        raise AssertionError("stark_verify rejected the generated proof: %s" % (why[0] if why else "?"))
    open(zkin_file, "w").write(js)
    return js


def timing_enable(on=True):
    _lib.lib().b200_timing_enable(1 if on else 0)


def timing_report():
    out = ctypes.c_void_p(); ln = ctypes.c_size_t()
    _lib.check(_lib.lib().b200_timing_report(ctypes.byref(out), ctypes.byref(ln)))
    return json.loads(_lib.take_string(out, ln))


# ---- compressor12 exec phase (recursion/src/compressor12/compressor12_exec.rs) without the `.cm` file round trip ------------------
def compressor12_exec(exec_vec, witness, n_rows, device_out_ptr=None):
    """exec_vec: the `.exec` vector (list / array of u64), witness: the circom witness as u64 values.  Returns the row-major
    (n_rows, 12) trace as a numpy array, or fills device memory at `device_out_ptr` (then returns None): the buffer
    `StarkProof.stark_gen(..., device_ptr=...)` takes."""
    e = np.ascontiguousarray(exec_vec, dtype=np.uint64); w = np.ascontiguousarray(witness, dtype=np.uint64)
    L = _lib.lib()
    if device_out_ptr is not None:
        _lib.check(L.b200_c12_exec_dev(_ptr(e), e.size, _ptr(w), w.size, n_rows, ctypes.c_void_p(device_out_ptr)))
        return None
    out = np.zeros((n_rows, 12), dtype=np.uint64)
    _lib.check(L.b200_c12_exec(_ptr(e), e.size, _ptr(w), w.size, n_rows, _ptr(out)))
    return out


def load_pols_dev(path, n_rows, n_cols, device_out_ptr):
    """`PolsArray::load` (starky/src/polsarray.rs:137-217) of a `.cm` / `.const` file straight into device memory"""
    _lib.check(_lib.lib().b200_pols_load_dev(path.encode(), n_rows, n_cols, ctypes.c_void_p(device_out_ptr)))
