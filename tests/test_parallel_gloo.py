"""CPU, world_size 2, gloo: the host-side sharding logic of the multi-GPU path (ray blocks + pixel
all-gather) reassembles exactly the unsharded frame."""
import os
import socket
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multiply_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_render(uv, P):
    """Deterministic per-ray 'render' so that sharded == unsharded can be asserted exactly."""
    x, y = uv[0, :, 0], uv[0, :, 1]
    rgb = torch.stack([x * 1e-3, y * 1e-3, (x + y) * 5e-4], 1)
    return {"rgb_values": rgb, "fg_rgb_values": rgb * 0.5, "normal_values": rgb - 0.1, "acc_map": x * 1e-3,
            "acc_person_list": torch.stack([x * (p + 1) * 1e-4 for p in range(P)], 1)}


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    inputs = {"uv": torch.rand(1, total, 2, generator=g) * 512, "pose": torch.eye(4)[None], "intrinsics": torch.eye(4)[None]}
    mine, (lo, hi) = parallel.shard_inputs(inputs, rank, world)
    out = _fake_render(mine["uv"], 2)
    full = parallel.gather_pixels(out, total)
    ref = _fake_render(inputs["uv"], 2)
    ok = all(torch.equal(full[k], ref[k]) for k in parallel.PIXEL_KEYS)
    q.put((rank, ok, lo, hi))
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    for total in (1, 7, 4096, 4097, 16384):
        for world in (1, 2, 3, 8):
            b = [parallel.shard_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_gather_equals_unsharded():
    world, total = 2, 1001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    assert sorted((lo, hi) for _, _, lo, hi in res) == [(0, 501), (501, 1001)]


# ---- person-sharded exchange: plan + all-to-all by ray block, world_size 2 and 3 ----
def test_exchange_plan_partitions_every_hit_list():
    g = torch.Generator().manual_seed(1)
    total = 1000
    hits = [torch.sort(torch.randperm(total, generator=g)[:k])[0] for k in (1, 37, 999, 1000)]
    hits.append(torch.zeros(1, dtype=torch.int64))       # the ray-0 substitute of an empty list
    assert [h.tolist() for h in parallel.normalize_hits([[], [3, 5]])] == [[0], [3, 5]]
    for world in (1, 2, 3, 8):
        plan = parallel.exchange_plan(hits, total, world)
        for h, pl in zip(hits, plan):
            assert pl[0][0] == 0 and pl[-1][1] == h.numel()
            for b, (lo, hi) in enumerate(pl):
                blo, bhi = parallel.shard_bounds(total, b, world)
                assert all(blo <= int(r) < bhi for r in h[lo:hi])
                if b + 1 < world:
                    assert hi == pl[b + 1][0]


def _xchg_worker(rank, world, port, total, P, width, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(5)
    hits = [torch.sort(torch.randperm(total, generator=g)[: 50 + 40 * p])[0] for p in range(P)]
    full = [torch.rand(h.numel(), width, generator=g) for h in hits]        # what the owners would compute
    plan = parallel.exchange_plan(hits, total, world)
    rows = {p: full[p] for p in range(P) if parallel.person_owner(p, world) == rank}
    got = parallel.exchange_person_rows(rows, plan, width, rank, world, torch.device("cpu"))
    lo, hi = parallel.shard_bounds(total, rank, world)
    ok = True
    for p in range(P):
        sel = (hits[p] >= lo) & (hits[p] < hi)
        ok = ok and torch.equal(got[p], full[p][sel])
    q.put((rank, ok))
    dist.destroy_process_group()


def test_person_rows_exchange_gloo():
    for world, P in ((2, 2), (3, 5)):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_xchg_worker, args=(r, world, port, 777, P, 9, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
        assert all(ok for _, ok in res), res


# ---- PixelBuffer: outputs written straight into the gather buffer, one all_gather_into_tensor, no concatenation ----
def _pixbuf_worker(rank, world, port, R, P, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    inputs = {"uv": torch.rand(1, R * world, 2, generator=g) * 512}
    mine = inputs["uv"][:, rank * R:(rank + 1) * R]
    buf = parallel.PixelBuffer(R, P, torch.device("cpu"))
    out = _fake_render(mine, P)
    for k in parallel.PIXEL_KEYS:
        assert buf.views[k].is_contiguous() and buf.views[k].shape == out[k].shape
        buf.views[k].copy_(out[k])            # stands in for Renderer.render(out=buf.views)
    gathered = torch.empty(world, R * (10 + P))
    dist.all_gather([gathered[r] for r in range(world)], buf.flat)    # NCCL path: all_gather_into_tensor(gathered, buf.flat)
    frame = parallel.PixelBuffer.frame(gathered, world, R, P)
    ref = _fake_render(inputs["uv"], P)
    q.put((rank, all(torch.equal(frame[k], ref[k]) for k in parallel.PIXEL_KEYS)))
    dist.destroy_process_group()


def test_pixel_buffer_all_gather_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pixbuf_worker, args=(r, world, port, 333, 3, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


# ---- the exchange inside a SUB-group: P2POp peers must be global ranks (ranks 1 and 2 of a 3-rank world) ----
def _subgroup_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grp = dist.new_group([1, 2])
    ok = True
    if rank in (1, 2):
        gr, gw = dist.get_rank(grp), 2
        g = torch.Generator().manual_seed(5)
        total, P, width = 300, 3, 5
        hits = [torch.sort(torch.randperm(total, generator=g)[: 40 + 30 * p])[0] for p in range(P)]
        full = [torch.rand(h.numel(), width, generator=g) for h in hits]
        plan = parallel.exchange_plan(hits, total, gw)
        rows = {p: full[p] for p in range(P) if parallel.person_owner(p, gw) == gr}
        got = parallel.exchange_person_rows(rows, plan, width, gr, gw, torch.device("cpu"), grp)
        lo, hi = parallel.shard_bounds(total, gr, gw)
        for p in range(P):
            sel = (hits[p] >= lo) & (hits[p] < hi)
            ok = ok and torch.equal(got[p], full[p][sel])
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


def test_person_rows_exchange_in_subgroup_gloo():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
