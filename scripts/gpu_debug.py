import os, sys, numpy as np, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiply_b200 import engine, scene as S
eng = os.environ.get("MP_ENGINE", "simt")
engine.set_engine(eng)
def ma(a, b): return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
for name, Sn, R, region in (("forward_S64_R48", 64, 48, "boxes"), ("forward_S16_R96", 16, 96, "image")):
    g = np.load(f"tests/golden/{name}.npz")
    sc = S.make_scene(P=2, S=Sn, seed=42)
    inp = S.make_rays(sc, R, seed=1234, region=region)
    hits = S.make_hit_lists(sc, inp)
    r = engine.Renderer(sc)
    o = r.render(inp, hits, debug=True)
    torch.cuda.synchronize()
    print(name, "trips", o["trips"].cpu().numpy(), g["trips"])
    for k in ("rgb_values", "fg_rgb_values", "normal_values", "acc_map", "acc_person_list", "bg_T"):
        print("  ", k, ma(o[k].cpu().numpy(), g[k]))
    bg = (o["rgb_values"] - (o["fg_rgb_values"] - o["bg_T"][:, None])).cpu().numpy() / np.maximum(o["bg_T"].cpu().numpy()[:, None], 1e-9)
    m = g["bg_T"] > 0.05
    print("   bg_rgb (where bg_T>0.05)", ma(bg[m], g["bg_rgb"][m]))
    for p in range(2):
        z = o[f"z_vals_{p}"].cpu().numpy()
        print("   p", p, "z", ma(z[:, :-1], g[f"z_vals_{p}"]), "sdf", ma(o[f"sdf_{p}"].cpu().numpy(), g[f"sdf_{p}"]),
              "rgb", ma(o[f"rgb_{p}"].cpu().numpy() * (g[f"sdf_{p}"] != 4.0)[..., None], g[f"rgb_{p}"] * (g[f"sdf_{p}"] != 4.0)[..., None]),
              "nrm", ma(o[f"normals_{p}"].cpu().numpy() * (g[f"sdf_{p}"] != 4.0)[..., None], g[f"normals_{p}"] * (g[f"sdf_{p}"] != 4.0)[..., None]),
              "outl mism", int(((o[f"sdf_{p}"].cpu().numpy() == 4.0) != (g[f"sdf_{p}"] == 4.0)).sum()))
        dz = np.abs(z[:, :-1] - g[f"z_vals_{p}"]); i = np.unravel_index(dz.argmax(), dz.shape); print("     worst z at", i, z[i[0], max(0,i[1]-2):i[1]+3], g[f"z_vals_{p}"][i[0], max(0,i[1]-2):i[1]+3])
# timing at config-2 shape
sc = S.make_scene(P=2, S=128, seed=42)
inp = S.make_rays(sc, 4096, seed=1234, region="boxes")
hits = S.make_hit_lists(sc, inp)
r = engine.Renderer(sc)
for i in range(3):
    torch.cuda.synchronize(); t = time.time(); o = r.render(inp, hits); torch.cuda.synchronize(); print("config2", eng, "render s", time.time() - t)
