"""ctypes binding of libb200zk.so (the C-ABI of include/b200zk.h).  Fails loudly when the library is missing;
compute calls fail with B200_ERR_CUDA when no GPU is present -- there is no CPU fallback."""
import ctypes, os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200ZK_LIB", os.path.join(_HERE, "libb200zk.so"))   # override only for kernel experiments
_LIB = None

u64p = ctypes.POINTER(ctypes.c_uint64)
_SIG = {
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_version": (ctypes.c_char_p, []),
    "b200_free": (None, [ctypes.c_void_p]),
    "b200_device_count": (ctypes.c_int, []),
    "b200_set_device": (ctypes.c_int, [ctypes.c_int]),
    "b200_set_stream": (ctypes.c_int, [ctypes.c_void_p]),
    "b200_timing_enable": (ctypes.c_int, [ctypes.c_int]),
    "b200_timing_report": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "b200_kernel_launches": (ctypes.c_uint64, []),
    "b200_gl_ntt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]),
    "b200_gl_intt": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]),
    "b200_gl_lde": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint]),
    "b200_gl_ntt_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int]),
    "b200_gl_lde_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint, ctypes.c_uint]),
    "b200_gl_poseidon": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_gl_linearhash": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_gl_merkle_n_nodes": (ctypes.c_size_t, [ctypes.c_size_t]),
    "b200_gl_merkelize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_gl_merkelize_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_big_poseidon": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_big_hash": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_big_linearhash": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_big_merkle_n_nodes": (ctypes.c_size_t, [ctypes.c_size_t]),
    "b200_big_merkelize": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_big_merkelize_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_setup_new": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "b200_setup_const_root": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "b200_setup_free": (None, [ctypes.c_void_p]),
    "b200_setup_shape": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "b200_setup_export": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "b200_setup_import": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]),
    "b200_setup_set_self_verify": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "b200_stark_verify": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_void_p)]),
    "b200_debug_step_program_source": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "b200_debug_jit_compile": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
    "b200_debug_msm_window": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint)]),
    "b200_debug_transcript_poseidon": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "b200_stark_gen": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "b200_msm_bn254_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bn254_g1_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_point_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "b200_msm": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bn254_g2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bn254_g2_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bls12381_g1": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bls12381_g1_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bls12381_g2": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_msm_bls12381_g2_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_c12_exec": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_c12_exec_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_pols_load_dev": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_groth16_pk_read": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]),
    "b200_groth16_pk_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p]),
    "b200_groth16_pk_free": (None, [ctypes.c_void_p]),
    "b200_groth16_prove": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t,
                                          ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_wtns_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "b200_msm_table_new": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "b200_msm_table_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_uint), ctypes.POINTER(ctypes.c_size_t)]),
    "b200_msm_table_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "b200_msm_table_set_partial_output": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "b200_msm_table_free": (None, [ctypes.c_void_p]),
    "b200_points_sum_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "b200_point_add": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_random_points_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]),
    "b200_bn254_g1_add": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "b200_bn254_g1_random_points_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64]),
    "b200_fr_fft": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int]),
    "b200_fr_fft_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint, ctypes.c_int]),
    "b200_groth16_h": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]),
    "b200_groth16_h_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p]),
    "b200_fib_trace_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint]),
    "b200_stark_gen_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
}
EXPORTS = sorted(_SIG)


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200zk error %d: %s" % (code, msg))
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libb200zk.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIG.items():
            f = getattr(L, name)        # raises AttributeError when a declared symbol is not exported
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise B200Error(rc, lib().b200_last_error().decode(errors="replace"))


def take_string(ptr, length):
    if not ptr.value:
        return ""
    s = (ctypes.string_at(ptr.value, length.value) if length is not None else ctypes.string_at(ptr.value)).decode()
    lib().b200_free(ptr)
    return s
