"""oracle/c12_exec.py pinned on the CPU: the `.exec` writer / reader pair restates the reference's own round-trip test
(recursion/src/compressor12/compressor12_exec.rs:120-150), and the fill is checked on a hand-computed example."""
from oracle import c12_exec as X


def test_write_and_read_exec_file_round_trip():
    # the reference's test data: no adds, a 12 x 3 signal map
    s_map = [[1, 2, 4], [2, 3, 42], [1, 1, 3], [4, 5, 2], [3, 4, 5], [1, 2, 4], [2, 3, 42], [1, 1, 3], [4, 5, 2], [3, 4, 5], [3, 4, 5], [3, 4, 5]]
    buf = X.write_exec([], s_map)
    adds_len, rows, adds, m = X.read_exec(buf)
    assert (adds_len, rows, adds) == (0, 3, [])
    assert m[:12] == [c[0] for c in s_map] and m[24:] == [c[2] for c in s_map]         # row-major: buff[2 + 12 i + c] = s_map[c][i]


def test_fill_by_hand():
    P = X.P
    w = [1, 10, 20, 30]
    # two PlonkAdd rows, the second uses the first: w4 = 2 w1 + 3 w2 = 80 ; w5 = (p - 1) w4 + 1 w3 = 30 - 80 mod p
    buf = X.write_exec([(1, 2, 2, 3), (4, 3, P - 1, 1)], [[1, 4], [5, 0]] + [[0, 2]] * 10)
    assert X.to_raw(1) == (1 << 64) % P                        # the raw form of ONE is R: what `adds[i].2.into()` stores for a coefficient of 1
    out = X.exec_fill(buf, w, 4)
    assert out.shape == (4, 12)
    assert [int(x) for x in out[0]] == [10, (30 - 80) % P] + [0] * 10
    assert [int(x) for x in out[1]] == [80, 0] + [20] * 10
    assert not out[2:].any()
