"""Ray-block sharding over the GPUs of one box.

Rays are independent except for the sampler's batch-global convergence flag, so the path shards with
no data-path collective: rank g renders the contiguous block [g*R/G, (g+1)*R/G) and one all-gather
assembles the pixel records.  Parity is defined per shard (a shard is exactly a smaller batch of
the reference)."""
import torch
import torch.distributed as dist

PIXEL_KEYS = ("rgb_values", "fg_rgb_values", "normal_values", "acc_map", "acc_person_list")


def shard_bounds(total, rank, world):
    """Contiguous, balanced ray blocks: the first (total % world) ranks get one extra ray."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(inputs, rank, world):
    lo, hi = shard_bounds(inputs["uv"].shape[1], rank, world)
    out = dict(inputs)
    out["uv"] = inputs["uv"][:, lo:hi].contiguous()
    return out, (lo, hi)


def pack_pixels(out):
    """[R_local, 3+3+3+1+P] record per ray."""
    return torch.cat([out["rgb_values"], out["fg_rgb_values"], out["normal_values"], out["acc_map"][:, None],
                      out["acc_person_list"]], dim=1).contiguous()


def unpack_pixels(rec, P):
    return {"rgb_values": rec[:, 0:3], "fg_rgb_values": rec[:, 3:6], "normal_values": rec[:, 6:9],
            "acc_map": rec[:, 9], "acc_person_list": rec[:, 10:10 + P]}


def gather_pixels(out, total, group=None):
    """All-gather the per-rank pixel records into the full frame (every rank gets it).  Blocks may differ by
    one ray, so records are padded to the largest block for the collective."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rec = pack_pixels(out)
    P = out["acc_person_list"].shape[1]
    if world == 1:
        return unpack_pixels(rec, P)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(mx, rec.shape[1], device=rec.device, dtype=rec.dtype)
    pad[: rec.shape[0]] = rec
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    full = torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
    return unpack_pixels(full, P)
