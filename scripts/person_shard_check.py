"""torchrun --nproc-per-node N scripts/person_shard_check.py : person-sharded frame == single-GPU frame, bit for bit,
plus the time of one person-sharded step (config 5 shape: 6 persons by default)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from multiply_b200 import engine, parallel, scene as S
P = int(os.environ.get("PERSONS", "6"))
R = int(os.environ.get("RAYS", "4096"))
Sn = int(os.environ.get("SAMPLES", "128"))
engine.set_engine("tc")
sc = S.make_scene(P=P, S=Sn, seed=42)
inp = S.make_rays(sc, R, seed=1234, region="boxes")
hits = S.make_hit_lists(sc, inp)
r = parallel.PersonShardedRenderer(sc, device=dev)
out = r.render(inp, hits)
torch.cuda.synchronize()
dist.barrier()
t0 = time.time()
for _ in range(5):
    out = r.render(inp, hits)
torch.cuda.synchronize()
dist.barrier()
dt = (time.time() - t0) / 5
ok = True
if dist.get_rank() == 0:
    ref = engine.Renderer(sc, device=dev).render(inp, hits)
    torch.cuda.synchronize()
    ok = all(torch.equal(out[k], ref[k]) for k in parallel.PIXEL_KEYS)
    print("PERSON_SHARD world", dist.get_world_size(), "persons", P, "rays", R, "bit_identical", ok,
          "ms_per_frame %.2f" % (dt * 1e3), "rays_per_s %.0f" % (R / dt))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
