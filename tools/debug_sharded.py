"""2-GPU determinism probe for sharded.lde_merkle_sharded (run under torchrun)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from eigen_zkvm_b200 import sharded, _lib
import bench
L = _lib.lib(); _lib.check(L.b200_set_device(local))
be = sharded.GpuBackend()
nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W = int(sys.argv[2]) if len(sys.argv) > 2 else 256
lo, hi = sharded.column_shard(W, world, rank)
cols = bench.splitmix_cols(torch, 1 << nbits, lo, hi, W, 0xE16E7)
def h(t): return int(t.view(-1)[::max(1, t.numel() // 65536)].sum().item()) & 0xFFFFFFFFFFFF
res = []
for it in range(3):
    ext = be.lde(cols, hi - lo, nbits, nbits + 3)
    rows = (1 << (nbits + 3)) // world
    send = ext.view(hi - lo, world, rows).permute(1, 0, 2).contiguous()
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv.view(-1), send.view(-1))
    torch.cuda.synchronize()
    nodes = be.merkelize(recv.view(-1), W, rows)
    nodes2 = be.merkelize(recv.view(-1), W, rows)
    res.append((h(ext), h(send), h(recv), h(nodes), h(nodes2), [int(x) for x in nodes[-4:].cpu().numpy().view(np.uint64)][:2]))
    print(rank, it, res[-1], flush=True)
dist.barrier(); dist.destroy_process_group()
