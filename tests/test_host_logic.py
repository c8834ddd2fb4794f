"""CPU tests of the host-side logic: tfd sample-shape semantics, the un-split [..,6] layout detection, sharding
helpers, and the N>1 path on a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp

import oracle as O
from vae_mdl_b200 import dist as vdist
from vae_mdl_b200._noise import sample_shape_to_n, uniform_noise


def test_sample_shape_semantics():
    assert sample_shape_to_n(()) == (1, True)      # tfd: sample() -> _sample_n(1), leading dim dropped
    assert sample_shape_to_n([]) == (1, True)
    assert sample_shape_to_n(7) == (7, False)
    assert sample_shape_to_n([3]) == (3, False)
    with pytest.raises(ValueError):
        sample_shape_to_n((2, 3))


def test_uniform_noise_range():
    u = uniform_noise((10000,), "cpu", torch.Generator().manual_seed(0))
    assert u.min().item() >= 1e-5 and u.max().item() <= 1 - 1e-5


def test_unsplit_layout_detection_needs_no_copy():
    from vae_mdl_b200.functional import _dl_layout
    both = torch.zeros(2, 3, 4, 4, 6)
    loc, ls = torch.split(both, 3, dim=-1)   # models/model03.py:91
    l2, s2, C, ld = _dl_layout(loc, ls)
    assert (C, ld) == (3, 6) and l2.data_ptr() == both.data_ptr() and s2.data_ptr() == both.data_ptr() + 12
    mu, lstd = torch.chunk(both, 2, dim=-1)
    assert _dl_layout(mu, lstd)[2:] == (3, 6)


def test_shard_bounds_cover_everything_once():
    for n in [0, 1, 7, 64, 26032]:
        for world in [1, 2, 3, 8]:
            seen = []
            for r in range(world):
                lo, hi = vdist.shard_bounds(n, r, world)
                seen += list(range(lo, hi))
                assert hi - lo in (n // world, n // world + 1)
            assert seen == list(range(n))
            rr = sorted(i for r in range(world) for i in vdist.round_robin(n, r, world))
            assert rr == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_images, S, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        table = torch.randn(n_images, S, generator=g, dtype=torch.float64) * 3  # log-weights of every (image, sample)

        def ll_chunk(i, s_lo, s_hi, out):
            out.copy_(table[i, s_lo:s_hi])

        ev = vdist.IwaeEvaluator(S, s_chunk=4, rank=rank, world=world, device="cpu")
        mean, llh = ev.run(n_images, ll_chunk, lambda lw: O.logmeanexp(lw, 0))
        want = O.logmeanexp(table.t(), 0).float()
        assert torch.allclose(llh, want, atol=1e-6), (llh, want)
        assert abs(mean.item() - want.mean().item()) < 1e-6

        # batch-sharded step: per-rank additive shares of the loss sum to the unsharded loss
        B = 5
        log_w = torch.randn(S, B, generator=g, dtype=torch.float64)
        lo, hi = vdist.shard_bounds(B, rank, world)

        def fake_step(params, x, extra, need_grad, b_total):
            lme = O.logmeanexp(extra, 0)
            return (-(lme.sum() / b_total)).reshape(1), extra, None

        loss, _, _ = vdist.sharded_modl_iwae_step(fake_step, None, None, log_w[:, lo:hi], B)
        want_loss = -O.logmeanexp(log_w, 0).mean()
        assert abs(loss.item() - want_loss.item()) < 1e-12
        # sample-sharded step (second row): every rank holds S / world samples of ALL images; the oracle stands
        # in for the kernels (ll_fn / bwd_fn are plain callables) and the result must equal the unsharded objective
        S2, B2, H2, W2, M2 = 6, 3, 2, 2, 3
        p_all = torch.randn(S2, B2, H2, W2, 10 * M2, generator=g, dtype=torch.float64)
        x01 = torch.randint(0, 256, (B2, H2, W2, 3), generator=g).double() / 255.0
        extra_all = torch.randn(S2, B2, generator=g, dtype=torch.float64)
        s_lo, s_hi = vdist.shard_bounds(S2, rank, world)

        def ll_fn(p, x):
            return O.modl_log_prob(p, x).sum((-1, -2, -3)).detach()

        def bwd_fn(p, x, g_img):
            q = p.clone().requires_grad_(True)
            (O.modl_log_prob(q, x).sum((-1, -2, -3)) * g_img.double()).sum().backward()
            return q.grad

        loss_s, _, dp_s = vdist.sample_sharded_iwae_step(ll_fn, bwd_fn, p_all[s_lo:s_hi], x01, extra_all[s_lo:s_hi], S2)
        q_all = p_all.clone().requires_grad_(True)
        full = -O.logmeanexp(O.modl_log_prob(q_all, x01).sum((-1, -2, -3)) + extra_all, 0).mean()
        full.backward()
        assert abs(loss_s.item() - full.item()) <= 1e-6 * abs(full.item())
        assert torch.allclose(dp_s, q_all.grad[s_lo:s_hi], rtol=1e-5, atol=1e-9)   # g_ll passes through float32
        torch.save(llh, os.path.join(out_dir, f"llh{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_images", [7, 8])
def test_sharded_evaluation_world_size_2_gloo(tmp_path, n_images):
    port = _free_port()
    tmp.spawn(_worker, args=(2, port, n_images, 10, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "llh0.pt"))
    b = torch.load(os.path.join(tmp_path, "llh1.pt"))
    assert torch.equal(a, b) and a.numel() == n_images


def test_fill_canvas_matches_the_reference_loop(tmp_path):
    """utils/utils.py:74-80 restated as its double loop; ``write_ppm`` round trip."""
    import numpy as np
    from vae_mdl_b200.utils import fill_canvas, normalize, write_ppm
    n, h, w, c = 3, 4, 5, 3
    img = torch.arange(n * n * h * w * c + h * w * c, dtype=torch.int64).reshape(n * n + 1, h, w, c).to(torch.uint8)
    want = np.empty([n * h, n * w, c], dtype=np.uint8)
    for i in range(n):
        for j in range(n):
            want[i * h:(i + 1) * h, j * w:(j + 1) * w, :] = img[i * n + j].numpy()
    got = fill_canvas(img, n, h, w, c)
    assert got.shape == (n * h, n * w, c) and np.array_equal(got.numpy(), want)
    assert torch.equal(normalize(img), img.float() / 255.0)
    path = str(tmp_path / "grid.ppm")
    write_ppm(path, got)
    raw = open(path, "rb").read()
    assert raw.startswith(b"P6\n15 12\n255\n") and raw[len(b"P6\n15 12\n255\n"):] == want.tobytes()
    with pytest.raises(ValueError):
        fill_canvas(img[:4], n, h, w, c)


def test_u8_to_unit_recipe_is_exact_for_every_byte():
    """csrc/common.cuh::u8_to_unit: k * fl(1/255) plus one FMA residual step equals the correctly rounded k / 255.f
    (utils/data.py:15-16) for all 256 bytes; the bare product does not.  float64 emulates the two FMAs exactly here."""
    import numpy as np
    r = np.float32(1.0) / np.float32(255.0)
    bare = 0
    for k in range(256):
        kf = np.float32(k)
        want = np.float32(kf / np.float32(255.0))
        q0 = np.float32(kf * r)
        rem = np.float32(np.float64(kf) - np.float64(q0) * 255.0)
        q1 = np.float32(np.float64(q0) + np.float64(rem) * np.float64(r))
        assert q1 == want, k
        bare += int(q0 != want)
    assert bare > 0


def _bench_timing_worker(rank, world, port, out_dir):
    """bench.py's multi-rank timing plumbing on a gloo group: host barrier, the device-side start gate (a tiny all-reduce on the
    timing stream; a plain all-reduce here), max over ranks."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench
        dev = torch.device("cpu")
        bench.barrier(world)
        bench.start_gate(world, dev)
        bench.start_gate(world, dev)          # (the gate tensor is reused)
        bench.start_gate(1, dev)              # no-op on one rank
        worst = bench.max_over_ranks(10.0 + rank, world, dev)
        torch.save(torch.tensor([worst]), os.path.join(out_dir, f"t{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_bench_start_gate_and_max_over_ranks_world_size_2_gloo(tmp_path):
    port = _free_port()
    tmp.spawn(_bench_timing_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(os.path.join(tmp_path, "t0.pt"))
    b = torch.load(os.path.join(tmp_path, "t1.pt"))
    assert a.item() == 11.0 and b.item() == 11.0      # every rank reports the slowest rank's time
