"""Device-side scene objects on top of the C ABI: packed fields, posed bodies, the fused renderer.

PyTorch is used for device memory and streams only; every computation is a call into
libmultiply_b200.so (see _lib.py).  No fallback path exists.
"""
import ctypes as C
import math
import torch

from . import _lib as L


def _dev(t, device):
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


def set_engine(name):
    """'tc' = tcgen05 split-fp16 tensor-core engine (default), 'simt' = fp32 validation engine."""
    L.check(L.lib().mp_set_engine({"simt": 0, "tc": 1}[name]), "mp_set_engine")


def set_precision(mode):
    """tcgen05 engine precision: 'parity' (default, three split terms), 'colour1' (single-term colour layers),
    'throughput' (single fp16 term everywhere; outside the 1e-4 gate)."""
    L.check(L.lib().mp_set_precision({"parity": 0, "colour1": 1, "throughput": 2}[mode]), "mp_set_precision")


def get_engine():
    return {0: "simt", 1: "tc"}[L.lib().mp_get_engine()]


def _stack(sd, n_layers, dev, keep):
    st = L.LinearStack()
    st.n_layers = n_layers
    for l in range(n_layers):
        if f"lin{l}.weight_v" in sd:
            v = _dev(sd[f"lin{l}.weight_v"], dev)
            g = _dev(sd[f"lin{l}.weight_g"], dev)
            keep += [v, g]
            st.weight_v[l] = v.data_ptr()
            st.weight_g[l] = g.data_ptr()
        else:
            v = _dev(sd[f"lin{l}.weight"], dev)
            keep.append(v)
            st.weight_v[l] = v.data_ptr()
            st.weight_g[l] = None
        b = _dev(sd[f"lin{l}.bias"], dev)
        keep.append(b)
        st.bias[l] = b.data_ptr()
        st.out_dim[l], st.in_dim[l] = v.shape
    return st


class Field:
    """Packed ImplicitNet + RenderingNet pair (mp_field_pack)."""

    def __init__(self, implicit_sd, render_sd, background=False, device="cuda"):
        lib = L.lib()
        self.device = torch.device(device)
        keep = []
        imp = L.ImplicitDesc()
        imp.lin = _stack(implicit_sd, 9, self.device, keep)
        imp.d_in = 4 if background else 3
        imp.multires = 10 if background else 6
        imp.cond_dim = 32 if background else 69
        imp.skip_layer = 4
        ren = L.RenderDesc()
        n_ren = len([k for k in render_sd if k.startswith("lin") and k.endswith(".bias") and "pose" not in k])
        ren.lin = _stack(render_sd, n_ren, self.device, keep)
        ren.mode = 1 if background else 0
        ren.multires_view = 4 if background else -1
        if not background:
            pw, pb = _dev(render_sd["lin_pose.weight"], self.device), _dev(render_sd["lin_pose.bias"], self.device)
            keep += [pw, pb]
            ren.lin_pose_weight, ren.lin_pose_bias = pw.data_ptr(), pb.data_ptr()
        nbytes = lib.mp_field_pack_bytes()
        self.storage = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        L.check(lib.mp_field_pack(C.byref(imp), C.byref(ren), int(background), self.storage.data_ptr(), nbytes,
                                  C.byref(h), L.stream_ptr()), "mp_field_pack")
        torch.cuda.current_stream().synchronize()   # raw parameter tensors may be released now
        self.handle = h
        self.background = background

    def set_cond(self, cond):
        c = _dev(cond.reshape(-1), self.device)
        L.check(L.lib().mp_field_set_cond(self.handle, c.data_ptr(), L.stream_ptr()), "mp_field_set_cond")
        self._cond = c

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.lib().mp_field_free(self.handle)
        except Exception:
            pass

    # operator-level entry points ------------------------------------------------------
    def implicit_forward(self, x, want_feat=True, want_grad=False):
        lib = L.lib()
        x = _dev(x, self.device)
        N = x.shape[0]
        sdf = torch.empty(N, device=self.device)
        feat = torch.empty(N, 256, device=self.device) if want_feat else None
        ws = torch.empty(lib.mp_mlp_workspace_bytes(N), dtype=torch.uint8, device=self.device)
        if want_grad:
            grad = torch.empty(N, 3, device=self.device)
            L.check(lib.mp_implicit_forward_grad(self.handle, x.data_ptr(), N, sdf.data_ptr(), L.ptr(feat),
                                                 grad.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()),
                    "mp_implicit_forward_grad")
            return sdf, feat, grad
        L.check(lib.mp_implicit_forward(self.handle, x.data_ptr(), N, sdf.data_ptr(), L.ptr(feat), ws.data_ptr(),
                                        ws.numel(), L.stream_ptr()), "mp_implicit_forward")
        return sdf, feat

    def sdf_grid(self, center, extent, res, pad=1.1):
        """Canonical SDF on the (res+1)^3 lattice of generate_mesh (lib/utils/mesh.py:78-105): returns [res+1]*3 fp32."""
        lib = L.lib()
        n1 = res + 1
        vals = torch.empty(n1, n1, n1, device=self.device)
        ws = torch.empty(lib.mp_sdf_grid_workspace_bytes(res), dtype=torch.uint8, device=self.device)
        c = (C.c_float * 3)(*[float(v) for v in center])
        L.check(lib.mp_sdf_grid(self.handle, c, float(extent), float(pad), int(res), vals.data_ptr(), ws.data_ptr(),
                                ws.numel(), L.stream_ptr()), "mp_sdf_grid")
        return vals

    def bg_forward(self, pts, view_dirs):
        """Background pair at given points (multiply.py:523-526): pts [N,4], view_dirs [N,3] -> (sdf [N], rgb [N,3])."""
        lib = L.lib()
        pts, view_dirs = _dev(pts, self.device), _dev(view_dirs, self.device)
        N = pts.shape[0]
        sdf = torch.empty(N, device=self.device)
        rgb = torch.empty(N, 3, device=self.device)
        ws = torch.empty(lib.mp_mlp_workspace_bytes(N), dtype=torch.uint8, device=self.device)
        L.check(lib.mp_bg_nets_forward(self.handle, pts.data_ptr(), view_dirs.data_ptr(), N, sdf.data_ptr(),
                                       rgb.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "mp_bg_nets_forward")
        return sdf, rgb

    def render_forward(self, points, normals, feat):
        lib = L.lib()
        points, normals, feat = (_dev(t, self.device) for t in (points, normals, feat))
        N = points.shape[0]
        rgb = torch.empty(N, 3, device=self.device)
        ws = torch.empty(lib.mp_mlp_workspace_bytes(N), dtype=torch.uint8, device=self.device)
        L.check(lib.mp_render_forward(self.handle, points.data_ptr(), normals.data_ptr(), feat.data_ptr(), N,
                                      rgb.data_ptr(), ws.data_ptr(), ws.numel(), L.stream_ptr()), "mp_render_forward")
        return rgb


class Body:
    """Canonical SMPL vertices + skinning weights (SMPLDeformer state) and the per-frame pose."""

    def __init__(self, verts_cano, weights, cano_cell=0.2, device="cuda"):
        lib = L.lib()
        self.device = torch.device(device)
        self.verts_c = _dev(verts_cano, self.device)
        self.weights = _dev(weights, self.device)
        V = self.verts_c.shape[0]
        nbytes = lib.mp_body_bytes(V)
        self.storage = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        L.check(lib.mp_body_create(self.verts_c.data_ptr(), self.weights.data_ptr(), V, float(cano_cell),
                                   self.storage.data_ptr(), nbytes, C.byref(h), L.stream_ptr()), "mp_body_create")
        self.handle = h
        self.V = V

    def set_pose(self, verts_posed, tfs):
        self.verts_p = _dev(verts_posed, self.device)
        self.tfs = _dev(tfs.reshape(24, 4, 4), self.device)
        L.check(L.lib().mp_body_set_pose(self.handle, self.verts_p.data_ptr(), self.tfs.data_ptr(), L.stream_ptr()),
                "mp_body_set_pose")

    def deform_inverse(self, x, exact_far=True):
        x = _dev(x, self.device)
        N = x.shape[0]
        xc = torch.empty(N, 3, device=self.device)
        out = torch.empty(N, dtype=torch.uint8, device=self.device)
        L.check(L.lib().mp_deform_inverse(self.handle, x.data_ptr(), N, xc.data_ptr(), out.data_ptr(),
                                          int(exact_far), L.stream_ptr()), "mp_deform_inverse")
        return xc, out.bool()

    def deform_broyden(self, x, max_steps=10, cvg_threshold=1e-5):
        """Root of forward_skinning(x_c) = x by Broyden's method from the closed-form inverse (row f4, not in the
        reference): returns dict(x_c, residual, converged, outlier, steps)."""
        x = _dev(x, self.device)
        N = x.shape[0]
        xc = torch.empty(N, 3, device=self.device)
        res = torch.empty(N, device=self.device)
        conv = torch.empty(N, dtype=torch.uint8, device=self.device)
        out = torch.empty(N, dtype=torch.uint8, device=self.device)
        steps = torch.empty(N, dtype=torch.int32, device=self.device)
        L.check(L.lib().mp_deform_broyden(self.handle, x.data_ptr(), N, int(max_steps), float(cvg_threshold),
                                          xc.data_ptr(), res.data_ptr(), conv.data_ptr(), out.data_ptr(),
                                          steps.data_ptr(), L.stream_ptr()), "mp_deform_broyden")
        return dict(x_c=xc, residual=res, converged=conv.bool(), outlier=out.bool(), steps=steps)

    def set_root_finder(self, max_steps, cvg_threshold=1e-5):
        """max_steps > 0: every inverse-deformer call on this body refines its non-outlier points with Broyden
        iterations (mp_body_set_root_finder); 0 restores the reference's closed-form inverse."""
        L.check(L.lib().mp_body_set_root_finder(self.handle, int(max_steps), float(cvg_threshold)),
                "mp_body_set_root_finder")

    def forward_jac(self, xc):
        xc = _dev(xc, self.device)
        N = xc.shape[0]
        xd = torch.empty(N, 3, device=self.device)
        J = torch.empty(N, 9, device=self.device)
        L.check(L.lib().mp_deform_forward_jac(self.handle, xc.data_ptr(), N, xd.data_ptr(), J.data_ptr(),
                                              L.stream_ptr()), "mp_deform_forward_jac")
        return xd, J

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                L.lib().mp_body_free(self.handle)
        except Exception:
            pass


def sampler_cfg(cfg, beta_param, beta_min=1e-4):
    c = L.SamplerCfg()
    c.scene_bounding_sphere = cfg["scene_bounding_sphere"]
    c.near = cfg.get("near", 0.0)
    c.N_samples = cfg["N_samples"]
    c.N_samples_eval = cfg["N_samples_eval"]
    c.N_samples_extra = cfg["N_samples_extra"]
    c.eps = cfg["eps"]
    c.beta_iters = cfg["beta_iters"]
    c.max_total_iters = cfg["max_total_iters"]
    c.add_tiny = cfg["add_tiny"]
    c.beta_param = beta_param
    c.beta_min = beta_min
    return c


class Renderer:
    """The fused eval forward (mp_render_rays) over a scene dict as produced by scene.make_scene
    (or assembled from a checkpoint by model.multiply.Multiply)."""

    def __init__(self, scene, device="cuda"):
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        device = self.device
        self.cfg = scene["cfg"]
        self.beta_param = float(scene["beta_param"])
        self.beta_min = float(scene.get("beta_min", 1e-4))
        self.P = len(scene["persons"])
        self.fields, self.bodies = [], []
        with torch.cuda.device(self.device):
            self._build(scene, device)
        self._ws = None
        self._status = None
        self.n = self.cfg["N_samples"] + self.cfg["N_samples_extra"] + 1

    def _build(self, scene, device):
        for person in scene["persons"]:
            f = Field(person["implicit"], person["render"], background=False, device=device)
            f.set_cond(person["cond"])
            scale = float(person.get("scale", 1.0))
            b = Body(person["verts_c"], person["weights"], cano_cell=0.1001 / max(scale, 1e-3), device=device)
            b.set_pose(person["verts_p"], person["tfs"])
            self.fields.append(f)
            self.bodies.append(b)
        self.bg = None
        if scene.get("bg_implicit") is not None:
            self.bg = Field(scene["bg_implicit"], scene["bg_render"], background=True, device=device)
            self.bg.set_cond(scene["frame_code"])

    def update_person(self, p, person):
        """New pose for person p (per-frame update): cond, posed vertices, bone transforms."""
        with torch.cuda.device(self.device):
            self.fields[p].set_cond(person["cond"])
            self.bodies[p].set_pose(person["verts_p"], person["tfs"])

    def check_status(self):
        """Raises if the last render saw a ray that misses the bounding sphere — where the reference prints
        'BOUNDING SPHERE PROBLEM!' and exits (rend_util.py:140-142).  Reads one device int (synchronises)."""
        if self._status is not None and int(self._status.item()) & 1:
            raise RuntimeError("BOUNDING SPHERE PROBLEM! (a camera ray misses the r=%g scene sphere)"
                               % self.cfg["scene_bounding_sphere"])

    def render(self, inputs, hit_lists, debug=False, persons=None, check=False, out=None, train=None):
        """inputs: uv [1,R,2], pose [1,4,4], intrinsics [1,4,4] (CUDA or CPU tensors);
        hit_lists: per rendered person either an int64 tensor of ray ids (empty -> ray 0, multiply.py:262-263) or a
        pair (ids [R] int64 on the device, count [1] int32 on the device) as produced by ``ray_aabb_hits`` — the
        count then never visits the host;
        persons: indices of the persons to render (default: all; ``Multiply.forward(input, id=p)`` passes [p],
        multiply.py:244-247) — ``acc_person_list`` has one column per rendered person;
        check: read the bounding-sphere status flag after the call (synchronises) and raise like the reference;
        out: optional dict of preallocated contiguous output tensors (e.g. ``parallel.PixelBuffer.views``);
        train: training-mode VALUES (mp_train_t): dict(rng=[per rendered person the tabled draws of
        ``ErrorBoundSampler.draw_training_rng``], t_rand_bg=[R,32] or None) — stochastic sampling, no outlier clamp,
        jittered background depths; adds ``z_eik_{k}`` [R_k] per person.  No gradients.
        Returns the eval output dict of Multiply.forward (multiply.py:589-598)."""
        with torch.cuda.device(self.device):
            return self._render(inputs, hit_lists, debug, persons, check, out, train)

    def _render(self, inputs, hit_lists, debug, persons, check, out_bufs=None, train=None):
        lib = L.lib()
        dev = self.device
        uv = _dev(inputs["uv"].reshape(-1, 2), dev)
        pose = _dev(inputs["pose"].reshape(4, 4), dev)
        K = _dev(inputs["intrinsics"].reshape(4, 4), dev)
        R = uv.shape[0]
        plist = list(range(self.P)) if persons is None else [int(p) for p in persons]
        Pn = len(plist)
        assert len(hit_lists) == Pn, "one hit list per rendered person"
        sc = L.Scene()
        sc.sampler = sampler_cfg(self.cfg, self.beta_param, self.beta_min)
        sc.P = Pn
        hits = []
        dev_counts = False
        for k, p in enumerate(plist):
            h = hit_lists[k]
            cnt = None
            if isinstance(h, (tuple, list)):
                h, cnt = h
                assert h.is_cuda and cnt.is_cuda and h.dtype == torch.int64 and cnt.dtype == torch.int32
                dev_counts = True
            elif h.numel() == 0:
                h = torch.zeros(1, dtype=torch.int64)
            h = h.to(device=dev, dtype=torch.int64).contiguous()
            hits.append((h, cnt))
            sc.body[k] = self.bodies[p].handle
            sc.field[k] = self.fields[p].handle
            sc.hit_index[k] = h.data_ptr()
            sc.hit_count[k] = h.numel()
            sc.hit_count_dev[k] = cnt.data_ptr() if cnt is not None else None
        sc.bg_field = self.bg.handle if self.bg is not None else None
        keep_train = []
        z_eik = {}
        if train is not None:
            assert not dev_counts, "training mode needs host-side hit counts"
            tr = L.Train()
            for k in range(Pn):
                rs, kp = sampler_rng_struct(train["rng"][k], dev)
                keep_train += [rs, kp]
                tr.rng[k] = C.pointer(rs)
                z_eik[k] = torch.empty(hits[k][0].numel(), device=dev)
                tr.z_eik[k] = z_eik[k].data_ptr()
            tb = train.get("t_rand_bg")
            if tb is not None:
                tb = _dev(tb, dev)
                keep_train.append(tb)
                tr.t_rand_bg = tb.data_ptr()
            keep_train.append(tr)
            sc.train = C.pointer(tr)
        need = lib.mp_render_workspace_bytes(C.byref(sc), R)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        if self._status is None:
            self._status = torch.zeros(1, dtype=torch.int32, device=dev)
        out = L.RenderOut()
        shapes = {"rgb_values": (R, 3), "fg_rgb_values": (R, 3), "normal_values": (R, 3), "acc_map": (R,),
                  "acc_person_list": (R, Pn)}
        if out_bufs is not None:
            res = {k: out_bufs[k] for k in shapes}
            for k, shp in shapes.items():
                t = res[k]
                assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shp, k
        else:
            res = {k: torch.empty(*shp, device=dev) for k, shp in shapes.items()}
        for k, v in res.items():
            setattr(out, k, v.data_ptr())
        out.status = self._status.data_ptr()
        dbg = {}
        if debug:
            assert not dev_counts, "debug taps need host-side hit counts"
            n = self.n
            dbg["trips"] = torch.zeros(Pn, dtype=torch.int32, device=dev)
            dbg["bg_T"] = torch.empty(R, device=dev)
            out.trips = dbg["trips"].data_ptr()
            out.bg_T = dbg["bg_T"].data_ptr()
            for k in range(Pn):
                Rp = hits[k][0].numel()
                dbg[f"z_vals_{k}"] = torch.empty(Rp, n + 1, device=dev)
                dbg[f"sdf_{k}"] = torch.empty(Rp, n, device=dev)
                dbg[f"rgb_{k}"] = torch.empty(Rp, n, 3, device=dev)
                dbg[f"normals_{k}"] = torch.empty(Rp, n, 3, device=dev)
                out.z_vals[k] = dbg[f"z_vals_{k}"].data_ptr()
                out.sdf[k] = dbg[f"sdf_{k}"].data_ptr()
                out.rgb[k] = dbg[f"rgb_{k}"].data_ptr()
                out.normals[k] = dbg[f"normals_{k}"].data_ptr()
        L.check(lib.mp_render_rays(C.byref(sc), uv.data_ptr(), pose.data_ptr(), K.data_ptr(), R, C.byref(out),
                                   self._ws.data_ptr(), self._ws.numel(), L.stream_ptr()), "mp_render_rays")
        self._keep = (uv, pose, K, hits, keep_train)
        for k, v in z_eik.items():
            dbg[f"z_eik_{k}"] = v
        if check:
            self.check_status()
        res.update(dbg)
        return res


def sampler_rng_struct(rng, dev):
    """mp_sampler_rng_t from the tabled draws of ``ErrorBoundSampler.draw_training_rng`` (t_rand [R,E], u_final [R,S],
    extra_perm [T,T*E] int32, eik_idx [T,R] int32, t_rand_bg [T,R,32]); returns (struct, tensors to keep alive)."""
    keep = {"t_rand": rng["t_rand"].to(device=dev, dtype=torch.float32).contiguous(),
            "u_final": rng["u_final"].to(device=dev, dtype=torch.float32).contiguous(),
            "extra_perm": rng["extra_perm"].to(device=dev, dtype=torch.int32).contiguous(),
            "eik_idx": rng["eik_idx"].to(device=dev, dtype=torch.int32).contiguous(),
            "t_rand_bg": rng["t_rand_bg"].to(device=dev, dtype=torch.float32).contiguous()}
    r = L.SamplerRng()
    for k, v in keep.items():
        setattr(r, k, v.data_ptr())
    return r, keep


class GraphedRender:
    """One eval forward captured in a CUDA graph and replayed: the ~60 kernel launches, memsets and the stream fork/join
    of mp_render_rays become a single graph launch (the sampler trips and the MLP tile counts are read from device
    memory, so nothing in the launch configuration depends on the data).  The input tensors are static device buffers —
    write new rays / camera into ``uv`` / ``pose`` / ``intrinsics`` (and new hit ids / counts into the hit-list
    tensors given at capture) and ``replay()``; poses are updated as usual through ``Renderer.update_person``."""

    def __init__(self, renderer, inputs, hit_lists, persons=None):
        self.r = renderer
        dev = renderer.device
        self.inputs = {k: _dev(inputs[k], dev).clone() for k in ("uv", "pose", "intrinsics")}
        self.hits = [tuple(t.to(dev) for t in h) if isinstance(h, (tuple, list)) else h.to(dev).clone() for h in hit_lists]
        self.persons = persons
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):                      # warm-up outside the capture: lazy resources, workspace size
                    renderer.render(self.inputs, self.hits, persons=persons)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = renderer.render(self.inputs, self.hits, persons=persons)

    def replay(self):
        self.graph.replay()
        return self.out


def ray_aabb_hits(cam_loc, ray_dirs, verts, inflate=1.2):
    """Device-side culling against the x`inflate` axis-aligned box of `verts` [V,3] (multiply.py:208-214 uses trimesh's
    oriented box): returns (ids [R] int64, count [1] int32), both on the device, the list already
    finalised (empty -> ray 0, multiply.py:262-263).  No host synchronisation: feed the pair to ``Renderer.render``."""
    lib = L.lib()
    dev = cam_loc.device
    cam = cam_loc.detach().contiguous().float()
    d = ray_dirs.detach().contiguous().float()
    v = verts.detach().reshape(-1, 3).contiguous().float()
    R = cam.shape[0]
    with torch.cuda.device(dev):
        idx = torch.empty(R, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        box = torch.empty(8, dtype=torch.float64, device=dev)
        L.check(lib.mp_ray_aabb_hits(cam.data_ptr(), d.data_ptr(), R, v.data_ptr(), v.shape[0], float(inflate),
                                     idx.data_ptr(), cnt.data_ptr(), box.data_ptr(), L.stream_ptr()), "mp_ray_aabb_hits")
    return idx, cnt


def ray_box_hits(cam_loc, ray_dirs, center, half_extent, rot=None, device_count=False):
    """Device-side ray / box culling (mp_ray_box_hits): sorted int64 ray ids that hit the box.  One scalar
    device->host read for the count — the reference pays a full `.cpu()` round trip plus trimesh here
    (multiply.py:256).  ``device_count=True`` returns (ids [R], count [1]) with the count left on the device and the
    list finalised (empty -> ray 0, mp_hit_list_finalize), the form ``Renderer.render`` takes without a host read."""
    lib = L.lib()
    dev = cam_loc.device
    cam = cam_loc.detach().contiguous().float()
    d = ray_dirs.detach().contiguous().float()
    R = cam.shape[0]
    c = (C.c_double * 3)(*[float(v) for v in center])
    h = (C.c_double * 3)(*[float(v) for v in half_extent])
    with torch.cuda.device(dev):
        idx = torch.empty(R, dtype=torch.int64, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        rot_d = None
        if rot is not None:
            rot_d = torch.as_tensor(rot, dtype=torch.float64).reshape(9).to(dev).contiguous()
        L.check(lib.mp_ray_box_hits(cam.data_ptr(), d.data_ptr(), R, c, h, L.ptr(rot_d), idx.data_ptr(), cnt.data_ptr(),
                                    L.stream_ptr()), "mp_ray_box_hits")
        if device_count:
            L.check(lib.mp_hit_list_finalize(idx.data_ptr(), cnt.data_ptr(), L.stream_ptr()), "mp_hit_list_finalize")
            return idx, cnt
    return idx[: int(cnt.item())]
