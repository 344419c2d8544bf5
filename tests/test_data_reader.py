"""Data-directory reader (multiply_b200/utils/data.py) against the layout Hi4D.py:119-146 reads."""
import struct
import zlib

import numpy as np
import pytest
import torch

from multiply_b200.utils import data as D


def _rot(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _camera(rng):
    K = np.array([[1200 + 50 * rng.random(), 0.3 * rng.random(), 470 + 10 * rng.random()],
                  [0, 1190 + 50 * rng.random(), 630 + 10 * rng.random()], [0, 0, 1.0]])
    R = _rot(rng)
    c = rng.standard_normal(3) * 2
    world = np.eye(4)
    world[:3, :4] = K @ np.concatenate([R, (-R @ c)[:, None]], 1)
    return K, R, c, world


def _write_png(path, h, w):
    raw = b"".join(b"\x00" + b"\x00" * (3 * w) for _ in range(h))
    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xffffffff)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def _make_dir(tmp_path, F=4, P=2, hw=(6, 8), seed=0):
    rng = np.random.default_rng(seed)
    np.save(tmp_path / "mean_shape.npy", rng.standard_normal((P, 10)))
    np.save(tmp_path / "poses.npy", rng.standard_normal((F, P, 72)) * 0.2)
    np.save(tmp_path / "normalize_trans.npy", rng.standard_normal((F, P, 3)) * 0.3)
    np.save(tmp_path / "gender.npy", np.array(["male", "female"][:P]))
    cams, truth = {}, []
    scale_mat = np.diag([2.5, 2.5, 2.5, 1.0])
    scale_mat[:3, 3] = [0.1, -0.2, 0.3]
    for i in range(F):
        K, R, c, world = _camera(rng)
        cams["scale_mat_%d" % i] = scale_mat
        cams["world_mat_%d" % i] = world
        truth.append((K, R, c))
    np.savez(tmp_path / "cameras_normalize.npz", **cams)
    (tmp_path / "image").mkdir()
    for i in range(F):
        _write_png(tmp_path / "image" / ("%04d.png" % i), *hw)
    return truth, scale_mat


def test_projection_decomposition_recovers_camera():
    rng = np.random.default_rng(1)
    for _ in range(20):
        K, R, c, world = _camera(rng)
        intr, pose = D.load_K_Rt_from_P(world[:3, :4])
        assert np.allclose(intr[:3, :3], K, atol=1e-6 * 1200)
        assert np.allclose(pose[:3, :3], R.T, atol=1e-6)
        assert np.allclose(pose[:3, 3], c, atol=1e-5)


def test_projection_decomposition_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2)
    for _ in range(20):
        _, _, _, world = _camera(rng)
        P = (world @ np.diag([3.0, 3.0, 3.0, 1.0]))[:3, :4].astype(np.float32)
        out = cv2.decomposeProjectionMatrix(P)                      # what rend_util.py:29 calls
        K_ref = out[0] / out[0][2, 2]
        pose_ref = np.eye(4, dtype=np.float32)
        pose_ref[:3, :3] = out[1].T
        pose_ref[:3, 3] = (out[2][:3] / out[2][3])[:, 0]
        intr, pose = D.load_K_Rt_from_P(P)
        assert np.allclose(intr[:3, :3], K_ref, rtol=1e-4, atol=1e-3)
        assert np.allclose(pose, pose_ref, atol=2e-5)


def test_sequence_directory(tmp_path):
    truth, scale_mat = _make_dir(tmp_path)
    seq = D.SequenceData(str(tmp_path), start_frame=1, end_frame=4)
    assert len(seq) == 3 and seq.num_person == 2 and seq.img_size == (6, 8) and seq.total_pixels == 48
    assert seq.gender == ["male", "female"]
    assert abs(seq.scale - 1 / 2.5) < 1e-7
    inp = seq[1]                                                    # frame 2 of the directory
    assert inp["uv"].shape == (1, 48, 2)
    assert inp["uv"][0, 0].tolist() == [0.0, 0.0] and inp["uv"][0, 1].tolist() == [1.0, 0.0]   # (x, y), row-major
    assert inp["uv"][0, 8].tolist() == [0.0, 1.0]
    sp = inp["smpl_params"]
    assert sp.shape == (1, 2, 86)
    assert torch.allclose(sp[0, :, 0], torch.full((2,), 1 / 2.5))
    assert np.allclose(sp[0, :, 1:4].numpy(), np.load(tmp_path / "normalize_trans.npy")[2], atol=1e-6)
    assert np.allclose(sp[0, :, 4:76].numpy(), np.load(tmp_path / "poses.npy")[2], atol=1e-6)
    assert np.allclose(sp[0, :, 76:].numpy(), np.load(tmp_path / "mean_shape.npy"), atol=1e-6)
    assert torch.equal(inp["smpl_pose"], sp[..., 4:76]) and torch.equal(inp["smpl_trans"], sp[..., 1:4])
    # camera of the normalised scene: centre = scale_mat^-1 (c), same rotation, K scaled by nothing
    K, R, c = truth[2]
    c_n = np.linalg.solve(scale_mat, np.append(c, 1.0))[:3]
    assert np.allclose(inp["pose"][0, :3, 3].numpy(), c_n, atol=1e-4)
    assert np.allclose(inp["pose"][0, :3, :3].numpy(), R.T, atol=1e-5)
    assert np.allclose(inp["intrinsics"][0, :3, :3].numpy(), K, rtol=1e-4, atol=1e-2)
    assert np.allclose(inp["C"], c_n, atol=1e-4)


def test_png_size_rejects_other_files(tmp_path):
    p = tmp_path / "x.png"
    p.write_bytes(b"not a png at all, just some bytes....")
    with pytest.raises(ValueError):
        D.png_size(str(p))


def test_body_params_from_checkpoint(tmp_path):
    """The optimised SMPL tables of a Lightning checkpoint (body_model_list.{p}.{name}.weight, multiply_model.py:81-92,
    body_model_params.py:5-50) override the directory's poses the way the opt_smpl branch does (:163-170)."""
    _make_dir(tmp_path, F=4, P=2)
    seq = D.SequenceData(str(tmp_path))
    g = torch.Generator().manual_seed(0)
    sd = {"model.density.beta": torch.tensor(0.1)}
    for p in range(2):
        sd["body_model_list.%d.betas.weight" % p] = torch.randn(1, 10, generator=g)
        sd["body_model_list.%d.global_orient.weight" % p] = torch.randn(4, 3, generator=g)
        sd["body_model_list.%d.body_pose.weight" % p] = torch.randn(4, 69, generator=g)
        sd["body_model_list.%d.transl.weight" % p] = torch.randn(4, 3, generator=g)
    bp = D.body_params_from_state_dict(sd)
    assert len(bp) == 2 and bp[1]["body_pose"].shape == (4, 69)
    plain = seq.frame(2)
    opt = seq.frame(2, body_params=bp)
    assert opt["smpl_pose"].shape == (1, 2, 72) and opt["smpl_shape"].shape == (1, 2, 10) and opt["smpl_trans"].shape == (1, 2, 3)
    for p in range(2):
        assert torch.equal(opt["smpl_pose"][0, p, :3], sd["body_model_list.%d.global_orient.weight" % p][2])
        assert torch.equal(opt["smpl_pose"][0, p, 3:], sd["body_model_list.%d.body_pose.weight" % p][2])
        assert torch.equal(opt["smpl_trans"][0, p], sd["body_model_list.%d.transl.weight" % p][2])
        assert torch.equal(opt["smpl_shape"][0, p], sd["body_model_list.%d.betas.weight" % p][0])
    assert torch.equal(opt["smpl_params"], plain["smpl_params"]) and torch.equal(opt["uv"], plain["uv"])   # scale etc. unchanged
    assert not torch.equal(opt["smpl_pose"], plain["smpl_pose"])
    assert D.body_params_from_state_dict({"model.density.beta": torch.tensor(0.1)}) == []
    del sd["body_model_list.1.transl.weight"]
    with pytest.raises(KeyError):
        D.body_params_from_state_dict(sd)
