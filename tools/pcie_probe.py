"""Host <-> device copy ceilings of the box, per GPU and in aggregate, with N ranks copying at the same time (what bounds the
end-to-end arm `vaemdl_modl_iwae_step_host` at N > 1).  One process per GPU:

    python tools/pcie_probe.py                                   # N = 1
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py

Every rank owns pinned host buffers and times, with CUDA events between barriers (max over ranks):
  h2d / d2h            one direction, contiguous 64 MiB cudaMemcpyAsync chunks
  both                 H2D and D2H at the same time on two streams (the step's steady state)
  both_2d              the same bytes as cudaMemcpy2DAsync of S strided rows per chunk (the [S, B, ...] -> [S, cb, ...] cut the
                       host-buffer step makes when it slices the batch)
Rank 0 prints one JSON line (per-GPU GB/s = bytes of ONE rank / time, aggregate = all ranks)."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import barrier, bind_near_gpu, dist_setup, max_over_ranks

MIB = 1 << 20


def main():
    rank, world, local = dist_setup(int(os.environ.get("WORLD_SIZE", "1")))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    bound = None if os.environ.get("PROBE_NO_BIND") else bind_near_gpu(local)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMemcpy2DAsync.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                     ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    H2D, D2H = 1, 2
    total = 1024 * MIB
    chunk = 64 * MIB
    S = 16                                   # rows per strided chunk, as the step's [S, B, ...] tensors
    h_in = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(total, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    h_out.fill_(0)
    d_in = torch.empty(chunk * 3, dtype=torch.uint8, device=dev)
    d_out = torch.empty(chunk * 3, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    n_chunks = total // chunk

    def up(i, strided):
        k = i % 3
        if not strided:
            d_in[k * chunk:(k + 1) * chunk].copy_(h_in[i * chunk:(i + 1) * chunk], non_blocking=True)
        else:  # S rows of chunk / S bytes, host pitch = total / S (row i of every chunk sits in row-block i of the buffer)
            w = chunk // S
            rc = rt.cudaMemcpy2DAsync(d_in.data_ptr() + k * chunk, w, h_in.data_ptr() + i * w, total // S, w, S, H2D,
                                      torch.cuda.current_stream(dev).cuda_stream)
            assert rc == 0, rc

    def down(i, strided):
        k = i % 3
        if not strided:
            h_out[i * chunk:(i + 1) * chunk].copy_(d_out[k * chunk:(k + 1) * chunk], non_blocking=True)
        else:
            w = chunk // S
            rc = rt.cudaMemcpy2DAsync(h_out.data_ptr() + i * w, total // S, d_out.data_ptr() + k * chunk, w, w, S, D2H,
                                      torch.cuda.current_stream(dev).cuda_stream)
            assert rc == 0, rc

    def run(do_up, do_dn, strided, reps=3):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize(dev)
            barrier(world)
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(torch.cuda.current_stream(dev))
            s_up.wait_event(e0)
            s_dn.wait_event(e0)
            if do_up:
                with torch.cuda.stream(s_up):
                    for i in range(n_chunks):
                        up(i, strided)
                    e1.record(s_up)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    for i in range(n_chunks):
                        down(i, strided)
                    e2.record(s_dn)
            torch.cuda.synchronize(dev)
            t = max(e0.elapsed_time(e1) if do_up else 0.0, e0.elapsed_time(e2) if do_dn else 0.0)
            t = max_over_ranks(t, world, dev)
            best = t if best is None else min(best, t)
        return best * 1e-3

    out = {"n_gpus": world, "bytes_per_direction_per_gpu": total, "chunk_mib": chunk // MIB,
           "cpus_bound_near_gpu": bound, "host_cpus": os.cpu_count()}
    for name, (u, d, s2) in {"h2d": (1, 0, 0), "d2h": (0, 1, 0), "both": (1, 1, 0), "both_2d": (1, 1, 1)}.items():
        t = run(u, d, s2)
        out[name] = {"seconds": t, "per_gpu_GBs_per_direction": total / t / 1e9,
                     "aggregate_GBs_all_directions": world * total * (u + d) / t / 1e9}
    if rank == 0:
        try:
            nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        except OSError:
            nodes = []
        out["numa_nodes"] = len(nodes)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
