"""Two (or more) GPUs, NCCL: the importance samples of one batch split across ranks (dist.sample_sharded_iwae_step) give
the loss and the gradient slices of the unsharded step.  torchrun --nproc-per-node N tools/nccl_split_s_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_mdl_b200 import dist as vdist, functional as F
rank, world, local = vdist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
S, B, H, W, M = 4 * world, 3, 32, 32, 10
g = torch.Generator().manual_seed(1)
params = torch.randn(S, B, H, W, 10 * M, generator=g).to(dev)
x_u8 = torch.randint(0, 256, (B, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
ll = F.modl_log_likelihood(params, x_u8, dtype=torch.float64)
extra = (ll.mean(0, keepdim=True) - ll).float() + torch.randn(S, B, generator=g).to(dev)
ref = F.modl_iwae_step(params, x_u8, extra)                       # every rank computes the unsharded answer
lo, hi = vdist.shard_bounds(S, rank, world)
loss, lpxz, dp = vdist.sample_sharded_iwae_step(lambda p, x: F.modl_log_likelihood(p, x, dtype=torch.float64),
                                                lambda p, x, gi: F.modl_backward(p, x, g_image=gi),
                                                params[lo:hi].contiguous(), x_u8, extra[lo:hi].contiguous(), S)
err_loss = abs(loss.item() + ref[3].item()) / abs(ref[3].item())
err_g = ((dp - ref[5][lo:hi]).norm() / ref[5][lo:hi].norm()).item()
assert torch.equal(lpxz, ref[0][lo:hi]) and err_loss <= 1e-6 and err_g <= 1e-5, (err_loss, err_g)
print(f"rank {rank}/{world}: split-S step ok (loss rel {err_loss:.1e}, gradient slice rel {err_g:.1e})", flush=True)
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()
