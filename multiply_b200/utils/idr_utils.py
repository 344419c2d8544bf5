"""Chunked full-frame rendering: the behaviour of the reference's ``split_input`` / ``merge_output``
(/root/reference/code/lib/utils/idr_utils.py:3-29) and of the loop MultiplyModel.test_step builds on them
(multiply_model.py:1235-1270).

The reference chunks a frame because its forward holds autograd state for every sample; here the chunk size only bounds
the per-call workspace (16 384 rays x 3 persons x 385 samples = 1.4 GB).  Semantics kept: chunks are consecutive pixel
blocks of ``n_pixels`` (the last one shorter), every other entry of the input dict is shared by all chunks, ``None``
outputs are dropped, per-pixel scalars come back flat ``[B * total_pixels]`` and vectors as ``[B * total_pixels, C]``.
"""
import torch


def split_input(model_input, total_pixels, n_pixels=10000):
    """One shallow copy of ``model_input`` per block of pixels; only ``uv`` [B, total_pixels, 2] is sliced."""
    uv = model_input["uv"]
    chunks = []
    for start in range(0, total_pixels, n_pixels):
        stop = min(start + n_pixels, total_pixels)
        chunks.append({**model_input, "uv": uv[:, start:stop].contiguous()})
    return chunks


def _stack_pixels(parts, batch_size, total_pixels):
    """Per-chunk tensors of one output entry -> the full-frame tensor in pixel order."""
    scalar = parts[0].dim() == 1
    width = 1 if scalar else parts[0].shape[-1]
    frame = torch.cat([p.reshape(batch_size, -1, width) for p in parts], dim=1)
    return frame.reshape(batch_size * total_pixels) if scalar else frame.reshape(batch_size * total_pixels, -1)


def merge_output(res, total_pixels, batch_size):
    """Inverse of ``split_input`` on the output side: ``res`` is the list of per-chunk output dicts."""
    return {name: _stack_pixels([r[name] for r in res], batch_size, total_pixels)
            for name, first in res[0].items() if first is not None}


def render_full_frame(model, inputs, total_pixels, n_pixels=16384, id=-1):
    """``model(chunk, id)`` over consecutive pixel chunks, merged — what multiply_model.py:1235-1270 does around the
    scene model.  ``model`` is anything with the ``Multiply.forward`` signature (multiply_b200.model.multiply)."""
    outputs = [{k: (v.detach() if torch.is_tensor(v) else v) for k, v in model(chunk, id).items()}
               for chunk in split_input(inputs, total_pixels, n_pixels=n_pixels)]
    return merge_output(outputs, total_pixels, 1)
