"""compressor12 exec phase -- TEST INFRASTRUCTURE (numpy / python ints), restating
recursion/src/compressor12/compressor12_exec.rs:19-108 and the `.exec` layout of compressor12_setup.rs:51-83.

exec vector (the JSON array of u64 the reference stores): [adds_len, map_rows, adds (4 per row: index a, index b, raw Montgomery limb
of coefficient a, of coefficient b), s_map (12 per mapped row, row-major)].  `FGL::from_raw_repr` takes the limb as the internal
representation (field_gl.rs:337-356), i.e. the canonical coefficient is limb / R mod p with R = 2^64 mod p.
The reference ships no `.exec` fixture (starky/data/fib.exec is a starkinfo JSON despite its name), so the layout is pinned by the
writer/reader pair of the reference's own round-trip test (compressor12_exec.rs:120-150), restated in tests/test_oracle_c12_exec.py."""
import numpy as np

P = 0xFFFFFFFF00000001
R = (1 << 64) % P
R_INV = pow(R, P - 2, P)


def to_raw(coef):
    """canonical coefficient -> the raw Montgomery limb `write_exec_file` stores (adds[i].2.into())"""
    return coef % P * R % P


def write_exec(adds, s_map_cols):
    """adds: [(a, b, coef_a, coef_b)] with canonical coefficients; s_map_cols: 12 lists of equal length (s_map[c][i]).
    compressor12_setup.rs:51-83"""
    assert len(s_map_cols) == 12
    rows = len(s_map_cols[0])
    buf = [len(adds), rows]
    for a, b, ka, kb in adds:
        buf += [a, b, to_raw(ka), to_raw(kb)]
    for i in range(rows):
        for c in range(12):
            buf.append(s_map_cols[c][i])
    return buf


def read_exec(buf):
    """compressor12_exec.rs:94-108"""
    adds_len, rows = int(buf[0]), int(buf[1])
    rest = buf[2:]
    assert len(rest) == adds_len * 4 + rows * 12
    return adds_len, rows, rest[:adds_len * 4], rest[adds_len * 4:]


def exec_fill(buf, witness, n_rows):
    """compressor12_exec.rs:43-92: returns the committed polynomials as a row-major (n_rows, 12) uint64 array"""
    adds_len, rows, adds, s_map = read_exec([int(x) for x in buf])
    w = [int(x) % P for x in witness]
    for i in range(adds_len):
        ka = adds[4 * i + 2] * R_INV % P; kb = adds[4 * i + 3] * R_INV % P
        w.append((w[adds[4 * i]] * ka + w[adds[4 * i + 1]] * kb) % P)
    out = np.zeros((n_rows, 12), dtype=np.uint64)
    for i in range(rows):
        for c in range(12):
            s = s_map[12 * i + c]
            out[i, c] = w[s] if s else 0
    return out
