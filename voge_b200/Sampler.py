"""Inverse rendering: scatter image features onto Gaussians (reference VoGE/Sampler.py:5-94)."""
import torch

from . import _C


def _num_vert(frag, vert_index, n_vert):
    if n_vert is None:
        n_vert = frag.num_vertices if hasattr(frag, 'num_vertices') else vert_index.max() + 1
    return n_vert


def _check_index_range(n_vert, vert_index):
    """The reference's `assert n_vert > vert_index.max()` (Sampler.py:11, :29): a host round trip per call, kept for
    the same error behaviour -- except inside a CUDA-graph capture (voge_b200.graphs), where the host can not wait;
    there the kernels' own bounds check (indices outside [0, n_vert) are skipped) is the guard."""
    if _C.capture_state is None:
        assert n_vert > vert_index.max()


def sample_features(frag, image, n_vert=None):
    """feat[n,:] = sum_{(pix,k): idx=n} w * image[pix,:] ;  wsum[n] = sum w.
    Equivalent dense form (Documentation.md:94-100): W[pix, n] = weight scattered by index;
    feat = W^T @ image, wsum = W.sum(pixels).  -> (vert_feature (n,c), vert_sum_weight (n,))."""
    vert_weight, vert_index = frag.vert_weight, frag.vert_index
    n_vert = _num_vert(frag, vert_index, n_vert)
    assert image.device == vert_index.device
    _check_index_range(n_vert, vert_index)
    assert vert_weight.shape[0] == image.shape[0] and vert_weight.shape[1] == image.shape[1] \
        and vert_weight.shape[2] == image.shape[2]
    return _SampleVoGE.apply(image, vert_weight, vert_index, int(n_vert))


def scatter_max_weight(frag, n_vert=None):
    vert_weight, vert_index = frag.vert_weight, frag.vert_index
    n_vert = _num_vert(frag, vert_index, n_vert)
    _check_index_range(n_vert, vert_index)
    return _ScatterMax.apply(vert_weight, vert_index, int(n_vert))


class _SampleVoGE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, vert_weight, vert_index, num_vert):
        vert_feature, vert_sum_weight = _C.sample_voge(image, vert_weight, vert_index, num_vert)
        ctx.save_for_backward(image, vert_weight, vert_index)
        return vert_feature, vert_sum_weight

    @staticmethod
    def backward(ctx, grad_vert_feature, grad_vert_weight_sum):
        image, vert_weight, vert_index = ctx.saved_tensors
        grad_image, grad_weight = _C.sample_voge_backward(image, vert_weight, vert_index,
                                                          grad_vert_feature.contiguous(),
                                                          grad_vert_weight_sum.contiguous())
        return grad_image, grad_weight, None, None


class _ScatterMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vert_weight, vert_index, num_vert):
        out = _C.scatter_max(vert_weight, vert_index, num_vert)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad):
        return None, None, None
