"""compressor12 exec phase on the device (b200_c12_exec[_dev], b200_pols_load_dev) against oracle/c12_exec.py."""
import os, tempfile
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import starky
    return starky


def _case(n_w, n_adds, rows, seed):
    from oracle import c12_exec as X
    rng = np.random.default_rng(seed)
    w = rng.integers(0, 2**63, size=n_w, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n_w, dtype=np.uint64)      # some values >= p: reduced like FGL::from
    w[0] = 1
    adds = []
    for i in range(n_adds):
        hi = n_w + i          # a row may use any earlier signal, including earlier PlonkAdd results
        adds.append((int(rng.integers(0, hi)), int(rng.integers(0, hi)), int(rng.integers(0, X.P, dtype=np.uint64)), int(rng.integers(0, X.P, dtype=np.uint64))))
    total = n_w + n_adds
    cols = [[int(rng.integers(0, total)) if rng.random() > 0.2 else 0 for _ in range(rows)] for _ in range(12)]
    return X.write_exec(adds, cols), w


@pytest.mark.parametrize("n_w,n_adds,rows,n_rows", [(5, 0, 3, 4), (50, 40, 30, 32), (1000, 3000, 4000, 4096), (200, 100, 0, 8), (30000, 20000, 1 << 16, 1 << 16)])
def test_fill_matches_oracle(sk, n_w, n_adds, rows, n_rows):
    import torch
    from oracle import c12_exec as X
    buf, w = _case(n_w, n_adds, rows, n_w + rows)
    want = X.exec_fill(buf, w, n_rows)
    got = sk.compressor12_exec(buf, w, n_rows)
    assert (got == want).all()
    d = torch.empty(n_rows * 12, dtype=torch.int64, device="cuda")
    sk.compressor12_exec(buf, w, n_rows, device_out_ptr=d.data_ptr())
    assert (d.cpu().numpy().view(np.uint64).reshape(n_rows, 12) == want).all()


def test_errors(sk):
    from eigen_zkvm_b200 import _lib
    buf, w = _case(20, 5, 6, 1)
    for bad in (buf[:-1], buf + [0], [len(buf)] + buf[1:]):
        with pytest.raises(_lib.B200Error):
            sk.compressor12_exec(bad, w, 8)
    with pytest.raises(_lib.B200Error):
        sk.compressor12_exec(buf, w, 4)                       # more mapped rows than the polynomial degree
    b2 = list(buf); b2[2] = 10**6                             # PlonkAdd refers to a signal that does not exist
    with pytest.raises(_lib.B200Error):
        sk.compressor12_exec(b2, w, 8)
    b3 = list(buf); b3[-1] = 10**6                            # signal map out of range
    with pytest.raises(_lib.B200Error):
        sk.compressor12_exec(b3, w, 8)
    b4 = list(buf); b4[4] = 0xFFFFFFFF00000001                # raw coefficient that is not a field representation (from_raw_repr fails)
    with pytest.raises(_lib.B200Error):
        sk.compressor12_exec(b4, w, 8)


def test_pols_file_streams_to_the_device(sk, golden_dir):
    import torch
    from eigen_zkvm_b200 import _lib
    path = os.path.join(golden_dir, "fib.cm.gl")
    want = np.fromfile(path, dtype="<u8")
    d = torch.empty(want.size, dtype=torch.int64, device="cuda")
    sk.load_pols_dev(path, 1024, 2, d.data_ptr())
    assert (d.cpu().numpy().view(np.uint64) == want).all()
    with pytest.raises(_lib.B200Error):
        sk.load_pols_dev(path, 1024, 3, d.data_ptr())         # short file
    with pytest.raises(_lib.B200Error):
        sk.load_pols_dev(path, 512, 2, d.data_ptr())          # longer than declared
    # a larger file than one staging buffer (2 x 64 MiB + tail), and a non-canonical value
    big = np.arange((17 << 20) + 5, dtype=np.uint64)
    with tempfile.NamedTemporaryFile(suffix=".cm") as f:
        big.tofile(f.name)
        d2 = torch.empty(big.size, dtype=torch.int64, device="cuda")
        sk.load_pols_dev(f.name, big.size, 1, d2.data_ptr())
        assert (d2.cpu().numpy().view(np.uint64) == big).all()
        big[12345] = 0xFFFFFFFF00000001
        big.tofile(f.name)
        with pytest.raises(_lib.B200Error):
            sk.load_pols_dev(f.name, big.size, 1, d2.data_ptr())
