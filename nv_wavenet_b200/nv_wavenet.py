"""Python binding with the surface of the reference's `pytorch/nv_wavenet.py` (class NVWaveNet, enum Impl,
column_major), on top of the kept C-ABI `wavenet_infer` -- through ctypes instead of the THC-era pybind wrapper
(`pytorch/wavenet_infer_wrapper.cpp`) that no longer builds against torch >= 2.

    wavenet = NVWaveNet(**model.export_weights())
    samples = wavenet.infer(cond_input, Impl.PERSISTENT)      # torch.cuda.IntTensor [batch, samples]

Same constructor arguments, same shape asserts, same layouts as nv_wavenet.py:55-196.
"""
import ctypes as C

import torch

from . import _lib


def interleave_lists(a, b, c, d, e, f, g):
    return [x for t in zip(a, b, c, d, e, f, g) for x in t]


def column_major(x):
    """PyTorch tensors are row major: return a contiguous transpose (nv_wavenet.py:33-49)."""
    assert x.is_contiguous
    if len(x.size()) == 1:
        return x
    if len(x.size()) == 3:
        assert x.size(2) == 1
        x = torch.squeeze(x)
    if len(x.size()) == 2:
        return torch.t(x).contiguous()
    if len(x.size()) == 4:
        return x.permute(3, 2, 1, 0).contiguous()


class Impl:
    AUTO, SINGLE_BLOCK, DUAL_BLOCK, PERSISTENT = 0, 1, 2, 3


class NVWaveNet:
    def __init__(self, embedding_prev, embedding_curr, conv_out_weight, conv_end_weight, dilate_weights, dilate_biases,
                 max_dilation, res_weights, res_biases, skip_weights, skip_biases, use_embed_tanh):
        lib = _lib.lib()
        self._lib = lib
        self.R, self.S, self.A = lib.get_R(), lib.get_S(), lib.get_A()
        self.max_dilation = max_dilation
        self.use_embed_tanh = use_embed_tanh
        assert embedding_prev.size() == (self.A, self.R), \
            "embedding_prev: {} doesn't match compiled nv-wavenet size: {}".format(embedding_prev.size(), (self.A, self.R))
        self.embedding_prev = column_major(torch.t(embedding_prev))
        assert embedding_curr.size() == (self.A, self.R), \
            "embedding_curr: {} doesn't match compiled nv-wavenet size: {}".format(embedding_curr.size(), (self.A, self.R))
        self.embedding_curr = column_major(torch.t(embedding_curr))
        assert conv_out_weight.size()[:2] == (self.A, self.S), \
            "conv_out_weight: {} doesn't match compiled nv-wavenet size: {}".format(conv_out_weight.size()[:2], (self.A, self.S))
        self.conv_out = column_major(conv_out_weight)
        assert conv_end_weight.size()[:2] == (self.A, self.A), \
            "conv_end_weight: {} doesn't match compiled nv-wavenet size: {}".format(conv_end_weight.size()[:2], (self.A, self.A))
        self.conv_end = column_major(conv_end_weight)

        dilate_weights_prev, dilate_weights_curr = [], []
        for weight in dilate_weights:
            assert weight.size(2) == 2, "nv-wavenet only supports kernel_size 2"
            assert weight.size()[:2] == (2 * self.R, self.R), \
                "dilated weight: {} doesn't match compiled nv-wavenet size: {}".format(weight.size()[:2], (2 * self.R, self.R))
            dilate_weights_prev.append(column_major(weight[:, :, 0]))
            dilate_weights_curr.append(column_major(weight[:, :, 1]))
        for bias in dilate_biases:
            assert bias.size(0) == 2 * self.R
        for weight in res_weights:
            assert weight.size()[:2] == (self.R, self.R)
        for bias in res_biases:
            assert bias.size(0) == self.R
        for weight in skip_weights:
            assert weight.size()[:2] == (self.S, self.R)
        for bias in skip_biases:
            assert bias.size(0) == self.S
        dilate_biases = [column_major(b) for b in dilate_biases]
        res_weights = [column_major(w) for w in res_weights]
        res_biases = [column_major(b) for b in res_biases]
        skip_weights = [column_major(w) for w in skip_weights]
        skip_biases = [column_major(b) for b in skip_biases]
        # There's an extra residual layer that's not used (nv_wavenet.py:139-141)
        res_weights.append(torch.zeros(self.R, self.R, device=res_weights[0].device if res_weights else None))
        res_biases.append(torch.zeros(self.R, device=res_biases[0].device if res_biases else None))
        assert len(res_biases) == len(skip_biases) == len(dilate_biases) and \
            len(res_weights) == len(skip_weights) == len(dilate_weights), \
            "Number of layers is inconsistent for different parameter types."
        self.num_layers = len(res_biases)
        self.layers = interleave_lists(dilate_weights_prev, dilate_weights_curr, dilate_biases, res_weights, res_biases,
                                       skip_weights, skip_biases)

    def infer(self, cond_input, implementation):
        # cond_input is channels x batch x num_layers x samples (nv_wavenet.py:172-196)
        assert cond_input.size()[0:3:2] == (2 * self.R, self.num_layers), \
            "Inputs are channels x batch x num_layers x samples; got {}".format(cond_input.size())
        batch_size, sample_count = cond_input.size(1), cond_input.size(3)
        cond_input = column_major(cond_input).float()
        samples = torch.empty((batch_size, sample_count), dtype=torch.int32, device="cuda")
        keep = [t.float().contiguous() for t in self.layers]

        def arr(k):
            return (C.c_void_p * self.num_layers)(*[keep[7 * l + k].data_ptr() for l in range(self.num_layers)])

        tensors = [self.embedding_prev.float().contiguous(), self.embedding_curr.float().contiguous(),
                   self.conv_out.float().contiguous(), self.conv_end.float().contiguous()]
        self._lib.wavenet_infer(sample_count, batch_size, tensors[0].data_ptr(), tensors[1].data_ptr(), self.num_layers,
                                self.max_dilation, arr(0), arr(1), arr(2), arr(3), arr(4), arr(5), arr(6),
                                tensors[2].data_ptr(), tensors[3].data_ptr(), int(bool(self.use_embed_tanh)),
                                cond_input.data_ptr(), implementation, samples.data_ptr())
        return samples
