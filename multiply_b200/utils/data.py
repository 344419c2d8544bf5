"""Data-directory reader for the rendering path: the per-frame ``inputs`` dict of the reference's dataset
(/root/reference/code/lib/datasets/Hi4D.py:89-327) without its image / mask / SAM side (cv2, training losses).

A preprocessed sequence directory holds
    mean_shape.npy              [P, 10]       betas of each person                       (Hi4D.py:119)
    poses.npy                   [F, P, 72]    axis-angle pose per frame and person       (Hi4D.py:121)
    normalize_trans.npy         [F, P, 3]     translation in the normalised scene        (Hi4D.py:122)
    cameras_normalize.npz       scale_mat_i, world_mat_i  [4, 4] per frame               (Hi4D.py:126-128)
    gender.npy                  [P] strings   (read by multiply_model.py / the SMPL servers, optional here)
    image/*.png                 only the size of the first image is needed               (Hi4D.py:106)

``SequenceData(root, ...)[i]`` returns what ``Hi4DDataset.__getitem__`` returns for ``num_sample <= 0`` (all pixels,
Hi4D.py:305-318) restricted to the keys ``Multiply.forward`` and ``MultiplyModel.test_step`` read, already batched
([1, ...]) and with ``smpl_pose`` / ``smpl_shape`` / ``smpl_trans`` split out of ``smpl_params`` the way
multiply_model.py:182-184 does for a model without optimised poses.
"""
import glob
import os
import struct

import numpy as np
import torch


def decompose_projection(P):
    """K, R, t of a 3x4 projection matrix — what ``cv2.decomposeProjectionMatrix`` returns to
    ``rend_util.load_K_Rt_from_P`` (rend_util.py:29-32): P[:, :3] = K R with K upper triangular with a positive
    diagonal and R orthonormal (RQ decomposition), t the homogeneous camera centre (null vector of P)."""
    P = np.asarray(P, dtype=np.float64)
    M = P[:, :3]
    # RQ through QR of the row-reversed transpose: M = K R
    rev = np.eye(3)[::-1]
    q, r = np.linalg.qr((rev @ M).T)
    K = rev @ r.T @ rev
    R = rev @ q.T
    s = np.sign(np.diag(K))
    s[s == 0] = 1.0
    K = K * s[None, :]
    R = s[:, None] * R
    # camera centre: P c = 0
    _, _, vt = np.linalg.svd(P)
    c = vt[-1]
    return K, R, c


def load_K_Rt_from_P(P):
    """rend_util.py:21-42 for a given P: intrinsics [4,4] (K normalised by K[2,2]) and the camera-to-world pose [4,4]
    (rotation R^T, translation = camera centre)."""
    K, R, c = decompose_projection(np.asarray(P)[:3, :4])
    K = K / K[2, 2]
    intrinsics = np.eye(4)
    intrinsics[:3, :3] = K
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = R.T
    pose[:3, 3] = c[:3] / c[3]
    return intrinsics, pose


def png_size(path):
    """(height, width) from the IHDR chunk — the only thing Hi4D.py:106 needs the first image for."""
    with open(path, "rb") as f:
        head = f.read(24)
    if head[:8] != b"\x89PNG\r\n\x1a\n" or head[12:16] != b"IHDR":
        raise ValueError("%s is not a PNG file" % path)
    w, h = struct.unpack(">II", head[16:24])
    return int(h), int(w)


def body_params_from_state_dict(state_dict):
    """The per-person optimised SMPL parameters a Lightning checkpoint of the reference carries next to the scene model
    (multiply_model.py:38-44,81-92: ``body_model_list.{p}`` are ``BodyModelParams`` modules, lib/model/
    body_model_params.py:5-50, whose tables are ``nn.Embedding`` weights): returns a list over persons of
    ``dict(betas [1,10], global_orient [F,3], body_pose [F,69], transl [F,3])``.  Empty list if the checkpoint has none
    (a model trained with ``opt_smpl = False`` reads the poses from the data directory instead)."""
    persons = {}
    for key, value in state_dict.items():
        parts = key.split(".")
        if len(parts) == 4 and parts[0] == "body_model_list" and parts[3] == "weight":
            persons.setdefault(int(parts[1]), {})[parts[2]] = value.detach().float()
    out = []
    for p in sorted(persons):
        missing = {"betas", "global_orient", "body_pose", "transl"} - set(persons[p])
        if missing:
            raise KeyError("body_model_list.%d lacks %s" % (p, sorted(missing)))
        out.append(persons[p])
    if out and sorted(persons) != list(range(len(out))):
        raise KeyError("body_model_list indices are not contiguous: %s" % sorted(persons))
    return out


def apply_body_params(inputs, body_params, frame_idx):
    """multiply_model.py:163-170 (the ``opt_smpl`` branch): the optimised tables replace ``smpl_pose`` / ``smpl_shape`` /
    ``smpl_trans`` of the input dict for frame ``frame_idx`` (``betas`` has a single row shared by all frames,
    body_model_params.py:46-47).  Returns a new dict; tensors follow the device of ``inputs['smpl_params']``."""
    dev = inputs["smpl_params"].device
    i = int(frame_idx)
    row = lambda t: t[i:i + 1].to(dev)
    out = dict(inputs)
    out["smpl_trans"] = torch.stack([row(bp["transl"]) for bp in body_params], dim=1)
    out["smpl_shape"] = torch.stack([bp["betas"][:1].to(dev) for bp in body_params], dim=1)
    go = torch.stack([row(bp["global_orient"]) for bp in body_params], dim=1)
    bpose = torch.stack([row(bp["body_pose"]) for bp in body_params], dim=1)
    out["smpl_pose"] = torch.cat((go, bpose), dim=2)
    return out


class SequenceData:
    """Frames ``range(start_frame, end_frame)`` of a preprocessed sequence (Hi4D.py:93-146)."""

    def __init__(self, root, start_frame=0, end_frame=None, img_size=None, pixel_per_batch=16384):
        self.root = root
        self.shape = np.load(os.path.join(root, "mean_shape.npy")).astype(np.float32)
        self.num_person = self.shape.shape[0]
        poses = np.load(os.path.join(root, "poses.npy"))
        if end_frame is None:
            end_frame = poses.shape[0]
        self.indices = list(range(start_frame, end_frame))
        self.poses = poses[self.indices].astype(np.float32)
        self.trans = np.load(os.path.join(root, "normalize_trans.npy"))[self.indices].astype(np.float32)
        gender = os.path.join(root, "gender.npy")
        self.gender = list(np.load(gender)) if os.path.exists(gender) else None

        cams = np.load(os.path.join(root, "cameras_normalize.npz"))
        self.scale_mat_all = [cams["scale_mat_%d" % i].astype(np.float32) for i in self.indices]
        self.world_mat_all = [cams["world_mat_%d" % i].astype(np.float32) for i in self.indices]
        self.scale = 1.0 / self.scale_mat_all[0][0, 0]                       # Hi4D.py:130
        self.P, self.C, self.intrinsics_all, self.pose_all = [], [], [], []
        for scale_mat, world_mat in zip(self.scale_mat_all, self.world_mat_all):
            P = world_mat @ scale_mat                                          # Hi4D.py:137-146
            self.P.append(P)
            self.C.append(-np.linalg.solve(P[:3, :3], P[:3, 3]))
            K, pose = load_K_Rt_from_P(P[:3, :4])
            self.intrinsics_all.append(torch.from_numpy(K).float())
            self.pose_all.append(torch.from_numpy(pose).float())

        if img_size is None:
            imgs = sorted(glob.glob(os.path.join(root, "image", "*.png")))
            if not imgs:
                raise FileNotFoundError("no image/*.png under %s and no img_size given" % root)
            img_size = png_size(imgs[self.indices[0]])
        self.img_size = tuple(int(v) for v in img_size)                       # (H, W)
        self.total_pixels = int(np.prod(self.img_size))
        self.pixel_per_batch = pixel_per_batch

    def __len__(self):
        return len(self.indices)

    def pixel_grid(self):
        """Hi4D.py:254-255: pixel centres as (x, y), row-major over the image."""
        uv = np.mgrid[:self.img_size[0], :self.img_size[1]].astype(np.int32)
        return np.flip(uv, axis=0).copy().transpose(1, 2, 0).astype(np.float32)

    def smpl_params(self, idx):
        """Hi4D.py:257-262: [P, 86] = scale | translation | pose | betas."""
        sp = torch.zeros(self.num_person, 86)
        sp[:, 0] = float(self.scale)
        sp[:, 1:4] = torch.from_numpy(self.trans[idx])
        sp[:, 4:76] = torch.from_numpy(self.poses[idx])
        sp[:, 76:] = torch.from_numpy(self.shape)
        return sp

    def __getitem__(self, idx, device=None):
        sp = self.smpl_params(idx)[None]
        inputs = {
            "uv": torch.from_numpy(self.pixel_grid().reshape(-1, 2))[None],
            "P": self.P[idx], "C": self.C[idx],
            "intrinsics": self.intrinsics_all[idx][None],
            "pose": self.pose_all[idx][None],
            "smpl_params": sp,
            "smpl_pose": sp[..., 4:76], "smpl_shape": sp[..., 76:], "smpl_trans": sp[..., 1:4],   # multiply_model.py:182-184
            "idx": torch.tensor([idx]),
            "img_size": self.img_size,
        }
        if device is not None:
            inputs = {k: (v.to(device) if torch.is_tensor(v) else v) for k, v in inputs.items()}
        return inputs

    def frame(self, idx, device=None, body_params=None):
        """Input dict of frame ``idx``; with ``body_params`` (``body_params_from_state_dict`` of a checkpoint trained with
        ``opt_smpl``) the optimised poses / shapes / translations replace the directory's, as in
        multiply_model.py:163-170."""
        inputs = self.__getitem__(idx, device)
        return apply_body_params(inputs, body_params, idx) if body_params else inputs
