import os, sys, numpy as np, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiply_b200 import engine, scene as S
from oracle import port
def ma(a, b): return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
torch.set_num_threads(32)
for (Sn, R, seed) in ((32, 384, 77), (64, 256, 5)):
    sc = S.make_scene(P=2, S=Sn, seed=42)
    inp = S.make_rays(sc, R, seed=seed, region="boxes")
    hits = S.make_hit_lists(sc, inp)
    st = {}
    t = time.time(); ref = port.multiply_forward(sc, inp, hits, stats=st, return_samples=True); print("oracle s", time.time() - t)
    for eng in ("simt", "tc"):
        engine.set_engine(eng)
        o = engine.Renderer(sc).render(inp, hits, debug=True)
        torch.cuda.synchronize()
        print(Sn, R, eng, "trips", o["trips"].cpu().numpy(), st["trips"])
        for k in ("rgb_values", "fg_rgb_values", "normal_values", "acc_map"):
            e = np.abs(o[k].cpu().numpy() - ref[k].numpy()); e = e.reshape(e.shape[0], -1).max(1)
            print("   ", k, "max", e.max(), "n>1e-4", int((e > 1e-4).sum()), "rays", np.nonzero(e > 1e-4)[0][:8])
        for p in range(2):
            z = o[f"z_vals_{p}"].cpu().numpy()[:, :-1]; zr = ref["_z_vals"][p].numpy()
            sd = o[f"sdf_{p}"].cpu().numpy(); sr = ref["_sdf"][p].numpy()
            dz = np.abs(z - zr).max(1); ds = np.abs(sd - sr).max(1)
            print("    p", p, "z max", dz.max(), "rows z>1e-4", np.nonzero(dz > 1e-4)[0][:6], "sdf max", ds.max(), "rows", np.nonzero(ds > 1e-4)[0][:6],
                  "outl mism", int(((sd == 4.0) != (sr == 4.0)).sum()))
