import os, sys, numpy as np, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multiply_b200 import engine, scene as S
from oracle import port
def ma(a, b): return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max())
def P(*a):
    print(*a); sys.stdout.flush()
sc = S.make_scene(P=2, S=64, seed=42)
# ---------------- background isolation (simt) ----------------
engine.set_engine("simt")
g = np.load("tests/golden/bg_nets.npz")
fb = engine.Field(sc["bg_implicit"], sc["bg_render"], background=True)
fb.set_cond(sc["frame_code"])
sdf, feat = fb.implicit_forward(torch.from_numpy(g["x"]))
torch.cuda.synchronize()
P("bg implicit sdf", ma(sdf.cpu().numpy(), g["out"][:, 0]), "feat", ma(feat.cpu().numpy(), g["out"][:, 1:]))
inp = S.make_rays(sc, 48, seed=1234, region="boxes")
dirs, cam = port.get_camera_params(inp["uv"], inp["pose"], inp["intrinsics"])
R = dirs.shape[1]
cam = cam.unsqueeze(1).repeat(1, R, 1).reshape(-1, 3); dirs = dirs.reshape(-1, 3)
tb = torch.linspace(0., 1., steps=32)
z_bg = (torch.zeros(R, 1) * (1. - tb) + torch.ones(R, 1) * tb) * (1. / 3.0)
with torch.no_grad():
    ref_bg = port.background_rgb(dirs, cam, sc, z_bg)
from multiply_b200 import _lib as L
lib = L.lib()
d_dirs, d_cam = dirs.cuda().contiguous(), cam.cuda().contiguous()
out = torch.empty(R, 3, device="cuda")
ws = torch.empty(lib.mp_background_workspace_bytes(R), dtype=torch.uint8, device="cuda")
L.check(lib.mp_background(fb.handle, d_dirs.data_ptr(), d_cam.data_ptr(), R, 3.0, out.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()))
torch.cuda.synchronize()
P("mp_background vs oracle", ma(out.cpu().numpy(), ref_bg.numpy()))
# the pts the kernel produced sit at the start of the workspace: [R*32,4]
pts = ws[: R * 32 * 16].view(torch.float32).reshape(R * 32, 4).cpu()
zb = torch.flip(z_bg, dims=[-1])
ref_pts = port.depth2pts_outside(cam.unsqueeze(1).repeat(1, 32, 1), dirs.unsqueeze(1).repeat(1, 32, 1), zb, 3.0).reshape(-1, 4)
P("bg points vs oracle", ma(pts.numpy(), ref_pts.numpy()))
# ---------------- tcgen05 engine ----------------
if os.environ.get("SKIP_TC") != "1":
    engine.set_engine("tc")
    p0 = sc["persons"][0]
    f = engine.Field(p0["implicit"], p0["render"]); f.set_cond(p0["cond"])
    g = np.load("tests/golden/implicit_fg.npz")
    x = torch.from_numpy(g["x"])
    t = time.time(); sdf, _ = f.implicit_forward(x, want_feat=False); torch.cuda.synchronize()
    P("tc sdf-only err", ma(sdf.cpu().numpy(), g["out"][:, 0]), "time", time.time() - t)
    P("  first vals", sdf[:4].cpu().numpy(), g["out"][:4, 0])
    sdf, feat = f.implicit_forward(x); torch.cuda.synchronize()
    P("tc fwd sdf err", ma(sdf.cpu().numpy(), g["out"][:, 0]), "feat err", ma(feat.cpu().numpy(), g["out"][:, 1:]))
    gg = np.load("tests/golden/implicit_fg_grad.npz")
    sdf, feat, grad = f.implicit_forward(torch.from_numpy(gg["x"]), want_grad=True); torch.cuda.synchronize()
    P("tc grad err", ma(grad.cpu().numpy(), gg["grad"]))
    P("  grad first", grad[:2].cpu().numpy(), gg["grad"][:2])
