"""Mirror of /root/reference/code/lib/model/multiply.py: the ``Multiply`` scene model whose eval-mode
``forward`` (multiply.py:174-598) drops onto the fused C-ABI entry ``mp_render_rays``.

Differences forced by the offline environment (documented in DESIGN.md):
  * the SMPL body model files are licence-gated, so SMPL servers / deformers are injected
    (``smpl_server_list``: objects with ``forward(scale, transl, thetas, betas) -> dict(smpl_verts, smpl_tfs,
    smpl_weights)`` and canonical ``verts_c``); ``scene.SyntheticSMPLServer`` is the offline stand-in;
  * ray/box hit lists (trimesh on the host in the reference, multiply.py:208-214,256) are taken from
    ``input['index_ray_box_list']`` when present, else computed by a host slab test against the x1.2 box.
"""
import numpy as np
import torch
import torch.nn as nn

from .. import engine, scene as S
from .networks import ImplicitNet, RenderingNet, _get
from .density import LaplaceDensity, AbsDensity
from .ray_sampler import ErrorBoundSampler
from .deformer import SMPLDeformer
from . import rend_util


class Multiply(nn.Module):
    def __init__(self, opt, betas_path=None, smpl_server_list=None, num_person=None):
        super().__init__()
        if smpl_server_list is None:
            raise ValueError("SMPL model files are not redistributable: pass smpl_server_list")
        self.using_nerfacc = True
        P = len(smpl_server_list) if num_person is None else num_person
        self.smpl_server_list = list(smpl_server_list)
        self.foreground_implicit_network_list = nn.ModuleList(
            [ImplicitNet(_get(opt, "implicit_network")) for _ in range(P)])
        self.foreground_rendering_network_list = nn.ModuleList(
            [RenderingNet(_get(opt, "rendering_network")) for _ in range(P)])
        self.with_bkgd = _get(opt, "with_bkgd", True)
        self.bg_implicit_network = ImplicitNet(_get(opt, "bg_implicit_network"))
        self.bg_rendering_network = RenderingNet(_get(opt, "bg_rendering_network"))
        self.frame_latent_encoder = nn.Embedding(_get(opt, "num_training_frames"), _get(opt, "dim_frame_encoding"))
        self.deformer_list = [SMPLDeformer(smpl_verts=s.verts_c, smpl_weights=s.weights, scale=getattr(s, "scale", 1.0))
                              for s in self.smpl_server_list]
        self.sdf_bounding_sphere = 3.0                                   # multiply.py:85
        d = _get(opt, "density")
        self.density = LaplaceDensity(**(dict(d) if not isinstance(d, dict) else d))
        self.bg_density = AbsDensity()
        rs = _get(opt, "ray_sampler")
        rs = dict(rs) if not isinstance(rs, dict) else dict(rs)
        rs.pop("N_samples_inverse_sphere", None)
        self.ray_sampler = ErrorBoundSampler(self.sdf_bounding_sphere, inverse_sphere_bg=True, **rs)
        self._renderer = None
        self._key = None

    # ---- packed device state ---------------------------------------------------------------
    @property
    def field_list(self):
        return self._ensure_renderer(next(self.parameters()).device).fields

    def _scene_dict(self, persons):
        return dict(cfg=dict(self.ray_sampler.cfg, multires=6, bg_multires=10, bg_multires_view=4),
                    persons=persons,
                    bg_implicit={k: v.detach() for k, v in self.bg_implicit_network.state_dict().items()},
                    bg_render={k: v.detach() for k, v in self.bg_rendering_network.state_dict().items()},
                    frame_code=torch.zeros(1, 32), beta_param=float(self.density.beta.detach()))

    def _ensure_renderer(self, device, persons=None):
        key = (str(device), tuple(int(p._version) for p in self.parameters()))
        if self._renderer is None or self._key != key:
            if persons is None:
                persons = [self._person_dict(p, None) for p in range(len(self.smpl_server_list))]
            self._renderer = engine.Renderer(self._scene_dict(persons), device=device)
            self._key = key
        return self._renderer

    def _person_dict(self, p, smpl_out, cond=None):
        srv = self.smpl_server_list[p]
        d = dict(implicit={k: v.detach() for k, v in self.foreground_implicit_network_list[p].state_dict().items()},
                 render={k: v.detach() for k, v in self.foreground_rendering_network_list[p].state_dict().items()},
                 verts_c=srv.verts_c.reshape(-1, 3), weights=srv.weights.reshape(-1, 24),
                 scale=getattr(srv, "scale", 1.0))
        if smpl_out is None:
            smpl_out = srv.canonical_output()
        d["verts_p"] = smpl_out["smpl_verts"].reshape(-1, 3)
        d["tfs"] = smpl_out["smpl_tfs"].reshape(24, 4, 4)
        d["cond"] = cond if cond is not None else torch.zeros(1, 69)
        return d

    # ---- operator surface the sampler / callers use (multiply.py:137-151) --------------------------
    def sdf_func_with_smpl_deformer(self, x, cond, smpl_tfs, smpl_verts, person_id):
        """multiply.py:137-151: canonicalise x against person ``person_id``'s posed SMPL (nearest vertex, inverse
        LBS), evaluate the SDF network there, set outliers (> 0.1 from the body, deformer.py:49) to sdf = 4 in eval
        mode.  Returns (sdf [N,1], x_c [N,3], feature [N,256])."""
        x_c, outlier_mask = self.deformer_list[person_id].forward(x, smpl_tfs, return_weights=False, inverse=True,
                                                                  smpl_verts=smpl_verts)
        output = self.foreground_implicit_network_list[person_id](x_c, cond, person_id=person_id)[0]
        sdf = output[:, 0:1].clone()
        if not self.training:
            sdf[outlier_mask] = 4.0                                         # multiply.py:142-143
        if not self.with_bkgd and self.sdf_bounding_sphere > 0.0:
            raise NotImplementedError("with_bkgd=False (sphere clamp, multiply.py:145-148) is not on the shipped path")
        feature = output[:, 1:]
        return sdf, x_c, feature

    # ---- Multiply.forward, eval branch -------------------------------------------------------
    def forward(self, input, id=-1, cond_zero_shit=False, canonical_pose=False):
        if self.training:
            raise NotImplementedError("training-mode forward/backward is a 'next' row (SURVEY.md §8f-1); "
                                      "call .eval() — validation/test steps do (multiply_model.py:982,1624)")
        if id != -1 or canonical_pose:
            raise NotImplementedError("single-person (id) / canonical-pose rendering: next row")
        dev = input["uv"].device
        smpl_params, smpl_pose = input["smpl_params"], input["smpl_pose"]
        scale = smpl_params[:, :, 0]
        smpl_shape, smpl_trans = input["smpl_shape"], input["smpl_trans"]
        P = smpl_trans.shape[1]
        persons = []
        for i in range(P):
            out = self.smpl_server_list[i](scale[:, i], smpl_trans[:, i], smpl_pose[:, i], smpl_shape[:, i])
            cond_pose = smpl_pose[:, i, 3:] / np.pi                        # multiply.py:270
            persons.append(self._person_dict(i, out, cond_pose))
        r = self._ensure_renderer(dev, persons)
        for i in range(P):
            r.update_person(i, persons[i])
        if "image_id" in input:
            frame = self.frame_latent_encoder(input["image_id"])          # multiply.py:407-410
        elif input.get("idx") is not None:
            frame = self.frame_latent_encoder(input["idx"])
        else:
            frame = None
        if frame is not None and r.bg is not None:
            r.bg.set_cond(frame.detach())
        hits = input.get("index_ray_box_list")
        if hits is None:
            # multiply.py:208-214, :256-263: rays vs the person's box inflated by 1.2, on the device
            dirs, cam = rend_util.get_camera_params(input["uv"], input["pose"], input["intrinsics"])
            dirs = dirs[0]
            cam = cam.expand(dirs.shape[0], 3).contiguous()
            hits = []
            for i in range(P):
                v = persons[i]["verts_p"]
                lo, hi = v.min(0)[0], v.max(0)[0]
                hits.append(engine.ray_box_hits(cam, dirs, ((lo + hi) / 2).tolist(), ((hi - lo) / 2 * 1.2).tolist()))
        bg_saved = r.bg
        if frame is None:
            r.bg = None                                                    # white background, multiply.py:540-541
        try:
            out = r.render(input, hits)
        finally:
            r.bg = bg_saved
        return {k: out[k] for k in ("acc_map", "acc_person_list", "rgb_values", "fg_rgb_values", "normal_values")}
