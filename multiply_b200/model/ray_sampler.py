"""Mirror of /root/reference/code/lib/model/ray_sampler.py (ErrorBoundSampler, eval mode)."""
import ctypes as C
import torch

from .. import _lib as L
from .. import engine


class ErrorBoundSampler:
    """ray_sampler.py:45-230.  Same constructor; ``get_z_vals`` keeps the reference's signature and return
    shape ``((z_vals, z_vals_bg), z_samples_eik)`` and runs the whole Algorithm-1 loop on the device
    (mp_sample_rays) without a host synchronisation."""

    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=0.0):
        if not inverse_sphere_bg:
            raise NotImplementedError("the reference always builds the sampler with inverse_sphere_bg=True "
                                      "(multiply.py:92)")
        self.cfg = dict(scene_bounding_sphere=scene_bounding_sphere, near=near, N_samples=N_samples,
                        N_samples_eval=N_samples_eval, N_samples_extra=N_samples_extra, eps=eps,
                        beta_iters=beta_iters, max_total_iters=max_total_iters, add_tiny=add_tiny)
        self.scene_bounding_sphere = scene_bounding_sphere
        self.N_samples, self.N_samples_eval, self.N_samples_extra = N_samples, N_samples_eval, N_samples_extra
        self._ws = None
        self.last_trips = None

    def get_z_vals(self, ray_dirs, cam_loc, model, cond, smpl_tfs, eval_mode, smpl_verts, person_id):
        """model: object exposing ``density`` (LaplaceDensity), ``deformer_list`` and ``field_list`` —
        model.multiply.Multiply does."""
        if getattr(model, "training", False):
            raise NotImplementedError("training-mode (stochastic) sampling is a 'next' row (SURVEY.md §8f-1)")
        lib = L.lib()
        dev = ray_dirs.device
        R = ray_dirs.shape[0]
        c = engine.sampler_cfg(self.cfg, float(model.density.beta.detach()), float(model.density.beta_min))
        body = model.deformer_list[person_id].body(dev)
        body.set_pose(smpl_verts[0], smpl_tfs[0] if smpl_tfs.ndim == 4 else smpl_tfs)
        field = model.field_list[person_id]
        field.set_cond(cond["smpl"])
        nz = self.N_samples + self.N_samples_extra + 2
        z = torch.empty(R, nz, device=dev)
        z_bg = torch.empty(R, 32, device=dev)
        trips = torch.zeros(1, dtype=torch.int32, device=dev)
        need = lib.mp_sampler_workspace_bytes(C.byref(c), R)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        d = ray_dirs.detach().contiguous().float()
        o = cam_loc.detach().contiguous().float()
        L.check(lib.mp_sample_rays(C.byref(c), body.handle, field.handle, d.data_ptr(), o.data_ptr(), R, z.data_ptr(),
                                   z_bg.data_ptr(), trips.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                                   L.stream_ptr()), "mp_sample_rays")
        self.last_trips = trips
        # z_samples_eik only feeds the training-time eikonal term (ray_sampler.py:211-213)
        return (z, z_bg), z[:, :1]
