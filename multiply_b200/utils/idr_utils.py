"""Mirror of /root/reference/code/lib/utils/idr_utils.py:3-29 (split_input / merge_output) plus the chunked full-frame
loop of MultiplyModel.test_step (multiply_model.py:1235-1270) on top of them.

The reference splits a frame into ``n_pixels``-ray chunks because its forward holds autograd state for every sample;
here the chunk size only bounds the per-call workspace (16 384 rays x 3 persons x 385 samples = 1.4 GB)."""
import torch


def split_input(model_input, total_pixels, n_pixels=10000):
    """idr_utils.py:3-15: list of input dicts whose ``uv`` holds consecutive blocks of ``n_pixels`` pixels."""
    split = []
    for indx in torch.split(torch.arange(total_pixels, device=model_input["uv"].device), n_pixels, dim=0):
        data = model_input.copy()
        data["uv"] = torch.index_select(model_input["uv"], 1, indx)
        split.append(data)
    return split


def merge_output(res, total_pixels, batch_size):
    """idr_utils.py:17-29: concatenates the per-chunk output dicts along the pixel axis."""
    model_outputs = {}
    for entry in res[0]:
        if res[0][entry] is None:
            continue
        if len(res[0][entry].shape) == 1:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, 1) for r in res],
                                             1).reshape(batch_size * total_pixels)
        else:
            model_outputs[entry] = torch.cat([r[entry].reshape(batch_size, -1, r[entry].shape[-1]) for r in res],
                                             1).reshape(batch_size * total_pixels, -1)
    return model_outputs


def render_full_frame(model, inputs, total_pixels, n_pixels=16384, id=-1):
    """The loop of multiply_model.py:1235-1270: ``model(batch, id)`` over consecutive pixel chunks, merged.
    ``model`` is anything with the ``Multiply.forward`` signature (the mirror in multiply_b200.model.multiply)."""
    res = []
    for chunk in split_input(inputs, total_pixels, n_pixels=n_pixels):
        out = model(chunk, id)
        res.append({k: v.detach() for k, v in out.items()})
    return merge_output(res, total_pixels, 1)
