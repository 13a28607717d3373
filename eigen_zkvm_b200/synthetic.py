"""Synthetic circuits for benchmarks and tests of shapes the reference ships no fixture for: a WIDE
Fibonacci PIL with `pairs` independent Fibonacci column pairs (2 * pairs committed columns) and `n_consts` constant columns
(column 0 = ISLAST, the others are committed in the constant tree but unconstrained), i.e. the compressor12 shape
12 committed + 31 constant (recursion/src/compressor12/compressor12_pil.rs:49-81) for pairs = 6, n_consts = 31.
The constraints are the reference's own Fibonacci identities (starky/data/fib.pil.json.gl) replicated per pair, so proofs are
verifiable by `stark_verify`; the trace generator matches `fib.cm.gl` for pair 0 and uses other seeds for the rest."""
import numpy as np

P = 0xFFFFFFFF00000001


def wide_fib_pil(nbits, pairs, n_consts):
    n = 1 << nbits
    num = lambda v: {"op": "number", "deg": 0, "value": str(v)}
    cm = lambda i, nxt=False: {"op": "cm", "deg": 1, "id": i, "next": nxt}
    const0 = {"op": "const", "deg": 1, "id": 0, "next": False}
    sub = lambda a, b, d=1: {"op": "sub", "deg": d, "values": [a, b]}
    mul = lambda a, b: {"op": "mul", "deg": 2, "values": [a, b]}
    add = lambda a, b: {"op": "add", "deg": 1, "values": [a, b]}
    refs = {"Wide.ISLAST": {"type": "constP", "id": 0, "polDeg": n, "isArray": False}}
    for k in range(1, n_consts):
        refs["Wide.K%d" % k] = {"type": "constP", "id": k, "polDeg": n, "isArray": False}
    exprs, idents = [], []
    for p in range(pairs):
        a, b = 2 * p, 2 * p + 1
        refs["Wide.a%d" % p] = {"type": "cmP", "id": a, "polDeg": n, "isArray": False}
        refs["Wide.b%d" % p] = {"type": "cmP", "id": b, "polDeg": n, "isArray": False}
        exprs.append(sub(mul(sub(num(1), const0), sub(cm(a, True), cm(b))), num(0), 2))
        exprs.append(sub(mul(sub(num(1), const0), sub(cm(b, True), add(cm(a), cm(b)))), num(0), 2))
    exprs.append(sub(mul(const0, sub(cm(1), {"op": "public", "deg": 0, "id": 0})), num(0), 2))
    for e in range(len(exprs)):
        idents.append({"e": e, "fileName": "wide.pil", "line": e + 1})
    return {"nCommitments": 2 * pairs, "nQ": 0, "nIm": 0, "nConstants": n_consts,
            "publics": [{"polType": "cmP", "polId": 1, "idx": n - 1, "id": 0, "name": "out"}],
            "references": refs, "expressions": exprs, "polIdentities": idents,
            "plookupIdentities": [], "permutationIdentities": [], "connectionIdentities": []}


def _fib_cols(n, a0, b0):
    """(a_i, b_i) with a' = b, b' = a + b mod p, vectorised by 2x2 matrix doubling in python ints for the chunk heads and numpy for
    the inner runs (exact, mod p)."""
    a = np.zeros(n, dtype=np.uint64); b = np.zeros(n, dtype=np.uint64)
    x, y = a0 % P, b0 % P
    for i in range(n):
        a[i] = x; b[i] = y
        x, y = y, (x + y) % P
    return a, b


def wide_fib_trace(nbits, pairs, n_consts, seed=0xE16E7):
    """Returns (cm row-major N x 2 pairs, const row-major N x n_consts) as flat uint64 arrays.  For sizes up to ~2^16 (python loop);
    larger traces are generated on the device by the bench (b200_fib_trace_dev per pair is not available: see bench.py)."""
    n = 1 << nbits
    cm = np.zeros((n, 2 * pairs), dtype=np.uint64)
    for p in range(pairs):
        a, b = _fib_cols(n, 1 + 3 * p, 2 + 5 * p)
        cm[:, 2 * p] = a; cm[:, 2 * p + 1] = b
    const = np.zeros((n, n_consts), dtype=np.uint64)
    const[n - 1, 0] = 1
    if n_consts > 1:
        rng = np.random.default_rng(seed)
        const[:, 1:] = rng.integers(0, 1 << 62, size=(n, n_consts - 1), dtype=np.uint64)
    return cm.reshape(-1), const.reshape(-1)
