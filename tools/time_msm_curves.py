import sys, os, ctypes, torch, numpy as np
sys.path.insert(0, os.getcwd())
from eigen_zkvm_b200 import groth16 as g16, starky
for cid, logn in ((1, 20), (3, 20), (0, 22), (2, 22)):
    m = 1 << logn; pw = g16.point_words(cid)
    db = torch.empty(m * pw, dtype=torch.int64, device="cuda")
    g16.random_points_dev(db.data_ptr(), m, 0xB254, cid)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    ds = torch.randint(0, 2**62, (m * 4,), dtype=torch.int64, device="cuda", generator=g)
    r0 = g16.multiexp_dev(db.data_ptr(), ds.data_ptr(), m, cid)
    starky.timing_enable(True)
    for _ in range(3): r1 = g16.multiexp_dev(db.data_ptr(), ds.data_ptr(), m, cid)
    rows = starky.timing_report(); starky.timing_enable(False)
    print(g16.CURVE_NAMES[cid], logn, {r["name"]: round(r["ms"] / 3, 2) for r in rows}, int(r1[0]) & 0xffff)
