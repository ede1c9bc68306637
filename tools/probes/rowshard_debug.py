"""2-GPU debug of the row-sharded adjoint: torchrun --nproc-per-node 2 tools/probes/rowshard_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import perm_equiv_graph_neural_cdes_b200 as P
from perm_equiv_graph_neural_cdes_b200 import rowshard as RS
from oracle import reference_path as R

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n, h, B, T = 512, 64, 2, 4
g = torch.Generator().manual_seed(3)
A = torch.stack([torch.from_numpy(R.synthetic_graph_path(n, T, 3 + b)).float() for b in range(B)]).to(dev)
ts = torch.arange(T, dtype=torch.float32, device=dev)
y0 = torch.randn((B, n, h), generator=g).to(dev); gy = torch.randn((B, n, h), generator=g).to(dev)
vf = P.PermEquivGraphVectorField(h, h, h, 3, 0, n, key=5).to(dev)
r0, r1 = RS.row_range(n, rank, world)
ctl = RS.RowShardedControl(ts, A[:, :, r0:r1].contiguous(), A[:, :, :, r0:r1].transpose(-1, -2).contiguous(), h, 3, flags=vf.flags)
for dt0 in (1.0, 0.25):
    vf.zero_grad()
    y = y0[:, r0:r1].clone().requires_grad_(True)
    yT = RS.diffeqsolve_rowsharded(vf, ctl, y, 0.0, 1.0, dt0)
    (yT * gy[:, r0:r1]).sum().backward()
    torch.cuda.synchronize()
    gp = torch.cat([p.grad.reshape(-1) for p in vf.parameters()])
    vexp = ctl._bufs[2].view(torch.int32).cpu().tolist()
    print(f"rank {rank} dt0 {dt0}: yT nan {int(torch.isnan(yT).sum())} gy0 nan {int(torch.isnan(y.grad).sum())} gparams nan {int(torch.isnan(gp).sum())} epoch {ctl._epoch.value} vexp {vexp} flags {ctl._bufs[4].view(torch.int32).cpu().tolist()}", flush=True)
dist.destroy_process_group()
