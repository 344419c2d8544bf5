"""utils/idr_utils.py: chunking semantics of the reference's split_input / merge_output (idr_utils.py:3-29)."""
import torch

from multiply_b200.utils import idr_utils


def test_split_merge_roundtrip():
    g = torch.Generator().manual_seed(0)
    total, B = 23, 1
    inp = {"uv": torch.rand(B, total, 2, generator=g), "pose": torch.eye(4)[None], "idx": torch.tensor([3])}
    chunks = idr_utils.split_input(inp, total, n_pixels=10)
    assert [c["uv"].shape[1] for c in chunks] == [10, 10, 3]
    assert all(c["pose"] is inp["pose"] and c["idx"] is inp["idx"] for c in chunks)      # shared, not copied
    assert torch.equal(torch.cat([c["uv"] for c in chunks], 1), inp["uv"])
    assert inp["uv"].shape[1] == total                                                   # input dict untouched

    # a fake per-chunk "model": scalar, vector and None outputs
    def model(chunk, id=-1):
        uv = chunk["uv"][0]
        return {"acc_map": uv.sum(-1), "rgb_values": torch.cat([uv, uv[:, :1] * 2], -1), "skipped": None}

    res = [model(c) for c in chunks]
    out = idr_utils.merge_output(res, total, B)
    assert set(out) == {"acc_map", "rgb_values"}
    assert out["acc_map"].shape == (total,) and out["rgb_values"].shape == (total, 3)
    assert torch.equal(out["acc_map"], inp["uv"][0].sum(-1))
    assert torch.equal(out["rgb_values"][:, :2], inp["uv"][0])
    full = idr_utils.render_full_frame(model, inp, total, n_pixels=7)
    assert torch.equal(full["rgb_values"], out["rgb_values"]) and torch.equal(full["acc_map"], out["acc_map"])


def test_exact_multiple_and_single_chunk():
    inp = {"uv": torch.arange(16.0).reshape(1, 8, 2)}
    assert [c["uv"].shape[1] for c in idr_utils.split_input(inp, 8, n_pixels=4)] == [4, 4]
    assert [c["uv"].shape[1] for c in idr_utils.split_input(inp, 8, n_pixels=100)] == [8]
