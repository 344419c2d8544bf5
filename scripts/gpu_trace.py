"""Print the cycle stamps of one tile of CTA 0 of the tcgen05 chain (MP_TC_KNOBS=2), shade program of bench config 2.
The stamps are compiled in only with -DMP_TC_TRACE=1: `scripts/build_variant.sh trace -DMP_TC_TRACE=1`, then run this with
MP_LIB=multiply_b200/_variants/lib_trace.so."""
import os, ctypes as C
os.environ["MP_TC_KNOBS"] = os.environ.get("MP_TC_KNOBS", "2")
import torch
from multiply_b200 import _lib as L
from multiply_b200.scene import make_scene, make_rays, make_hit_lists
from multiply_b200.engine import Renderer

scene = make_scene(P=1, S=128, seed=0)
inputs = make_rays(scene, 4096, seed=1, region="boxes")
hits = make_hit_lists(scene, inputs)
r = Renderer(scene)
for _ in range(2):
    out = r.render(inputs, hits)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 4096)()
L.check(L.lib().mp_tc_trace_read(buf, 4096), "trace")
t = list(buf)
if not any(t):
    raise SystemExit("no stamps: this library was built without -DMP_TC_TRACE=1 (see the docstring)")
# the last launch is the shade chain of the single person (the background field is skipped for P=1? print both halves anyway)
fg = t[512:518]
print("final-grad step: reload_done chunks_done bar1 pairs_done bar2 normal_done:", [x - min(v for v in fg if v) if x else -1 for x in fg])
ex = t[520:536]
print("extra-in chunks (start, after extra loop) x4:", [x - min(v for v in ex if v) if x else -1 for x in ex])
for name, off in (("shade chain (20 steps)", 0), ("background / sdf-only chain", 1024)):
    seg = [x for x in t[off:off + 8 * 24] if x]
    if not seg:
        continue
    base = min(seg)
    print(name)
    print("step | wait_start acc_ready c0 c1 c2 c3 end || mma: kb0 kb1 kb2 kb3 commit   (cycles from first stamp)")
    for s in range(24):
        e = t[off + s * 8:off + s * 8 + 7]
        m = t[2048 + off + s * 8:2048 + off + s * 8 + 5]
        if not any(e):
            continue
        print("%2d | %s || %s" % (s, " ".join("%7d" % (x - base if x else -1) for x in e), " ".join("%7d" % (x - base if x else -1) for x in m)))
