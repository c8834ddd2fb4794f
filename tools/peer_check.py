"""Two or more GPUs: the ELBO shares of a batch split across ranks, exchanged over peer memory (vae_mdl_b200/peer.py:
NVLink P2P stores from the kernel that forms the share), equal the NCCL all-reduce and the unsharded step.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/peer_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as tdist
import vae_mdl_b200 as V
from vae_mdl_b200 import dist as vdist
from vae_mdl_b200.peer import ElboExchange
rank, world, local = vdist.init_from_env()
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
S, Bg, H, W, M = 5, 8 * world, 32, 32, 10
g = torch.Generator().manual_seed(1)
params = torch.randn(S, Bg, H, W, 10 * M, generator=g).to(dev)
x_u8 = torch.randint(0, 256, (Bg, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
extra = torch.randn(S, Bg, generator=g).to(dev)
ref_loss, _, ref_dp = V.modl_iwae_step(params, x_u8, extra)                 # every rank computes the unsharded answer
lo, hi = vdist.shard_bounds(Bg, rank, world)
shard = (params[:, lo:hi].contiguous(), x_u8[lo:hi].contiguous(), extra[:, lo:hi].contiguous())
ex = ElboExchange(dev)
for it in range(100):   # more steps than the ring holds
    loss_p, _, dp = vdist.sharded_modl_iwae_step(V.modl_iwae_step, *shard, b_total=Bg, exchange=ex)
loss_n, _, _ = vdist.sharded_modl_iwae_step(V.modl_iwae_step, *shard, b_total=Bg)
torch.cuda.synchronize()
err_p = abs(loss_p.item() - ref_loss.item()) / abs(ref_loss.item())
err_n = abs(loss_n.item() - ref_loss.item()) / abs(ref_loss.item())
err_g = ((dp - ref_dp[:, lo:hi]).norm() / ref_dp[:, lo:hi].norm()).item()
all_p = [torch.zeros(1, device=dev) for _ in range(world)]
tdist.all_gather(all_p, loss_p.reshape(1))
assert all(torch.equal(t, all_p[0]) for t in all_p), "the peer-memory sum differs between ranks"
assert err_p <= 1e-6 and err_n <= 1e-6 and err_g <= 1e-5, (err_p, err_n, err_g)
# time: step + peer-memory exchange vs step + NCCL all-reduce (us per step, max over ranks)
def timeit(fn, iters=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); tdist.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
    return t.item()
t_peer = timeit(lambda: vdist.sharded_modl_iwae_step(V.modl_iwae_step, *shard, b_total=Bg, exchange=ex))
t_nccl = timeit(lambda: vdist.sharded_modl_iwae_step(V.modl_iwae_step, *shard, b_total=Bg))
t_none = timeit(lambda: V.modl_iwae_step(shard[0], shard[1], shard[2], True, Bg))
print(f"rank {rank}/{world}: peer-memory ELBO exchange ok (loss rel {err_p:.1e}, nccl rel {err_n:.1e}, grad rel {err_g:.1e}); "
      f"us per step: no exchange {t_none:.1f}, peer memory {t_peer:.1f}, nccl all_reduce {t_nccl:.1f}", flush=True)
ex.close()
tdist.barrier()
tdist.destroy_process_group()
