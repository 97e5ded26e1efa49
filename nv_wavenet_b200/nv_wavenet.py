"""PyTorch-facing binding with the surface of the reference's `pytorch/nv_wavenet.py` -- class `NVWaveNet`, the `Impl`
constants and `column_major` -- implemented over the kept C-ABI `wavenet_infer` through ctypes (the reference's
THC-era pybind wrapper, `pytorch/wavenet_infer_wrapper.cpp`, no longer builds against torch >= 2).

    wavenet = NVWaveNet(**model.export_weights())
    samples = wavenet.infer(cond_input, Impl.PERSISTENT)      # int32 CUDA tensor [batch, samples]

Constructor arguments, accepted shapes, the appended unused residual layer and the memory layouts handed to the
kernel are those of pytorch/nv_wavenet.py:55-196; the code is organised around one table of expected shapes.
"""
import ctypes as C

import torch

from . import _lib


class Impl:
    """`implementation` argument of infer() (pytorch/nv_wavenet.py:51-54); one kernel family serves all of them here."""
    AUTO = 0
    SINGLE_BLOCK = 1
    DUAL_BLOCK = 2
    PERSISTENT = 3


def column_major(x):
    """Row-major torch tensor -> the column-major layout the kernel reads (pytorch/nv_wavenet.py:33-49):
    vectors unchanged, [M][K] and conv-style [M][K][1] matrices transposed, 4-D conditioning fully reversed."""
    nd = x.dim()
    if nd == 1:
        return x
    if nd == 3:
        if x.size(2) != 1:
            raise AssertionError("column_major: 3-D tensors must be convolution weights of kernel size 1")
        x, nd = x[:, :, 0], 2
    if nd == 2:
        return x.t().contiguous()
    if nd == 4:
        return x.permute(3, 2, 1, 0).contiguous()
    raise AssertionError(f"column_major: unsupported rank {nd}")


def _expect(name, tensor, shape):
    got = tuple(tensor.size())[:len(shape)]
    if got != tuple(shape):
        raise AssertionError(f"{name}: shape {got} does not match the compiled kernel's {tuple(shape)}")


class NVWaveNet:
    def __init__(self, embedding_prev, embedding_curr, conv_out_weight, conv_end_weight, dilate_weights, dilate_biases,
                 max_dilation, res_weights, res_biases, skip_weights, skip_biases, use_embed_tanh):
        self._lib = _lib.lib()
        R, S, A = self._lib.get_R(), self._lib.get_S(), self._lib.get_A()
        self.R, self.S, self.A = R, S, A
        self.max_dilation = max_dilation
        self.use_embed_tanh = use_embed_tanh

        # embeddings arrive [A][R] and are consumed as emb[a * R + r]: two transposes cancel, keep the values as they are
        _expect("embedding_prev", embedding_prev, (A, R))
        _expect("embedding_curr", embedding_curr, (A, R))
        self.embedding_prev = column_major(embedding_prev.t())
        self.embedding_curr = column_major(embedding_curr.t())
        _expect("conv_out_weight", conv_out_weight, (A, S))
        _expect("conv_end_weight", conv_end_weight, (A, A))
        self.conv_out = column_major(conv_out_weight)
        self.conv_end = column_major(conv_end_weight)

        n = len(dilate_weights)
        counts = {"dilate_biases": len(dilate_biases), "skip_weights": len(skip_weights), "skip_biases": len(skip_biases),
                  "res_weights": len(res_weights) + 1, "res_biases": len(res_biases) + 1}      # the last layer has no residual conv
        if any(c != n for c in counts.values()):
            raise AssertionError(f"Number of layers is inconsistent for different parameter types: dilate_weights {n}, {counts}")
        # the kernel still wants a residual matrix for the last layer: all zero (pytorch/nv_wavenet.py:139-141)
        like = dilate_weights[0] if n else embedding_prev
        res_weights = list(res_weights) + [torch.zeros(R, R, dtype=like.dtype, device=like.device)]
        res_biases = list(res_biases) + [torch.zeros(R, dtype=like.dtype, device=like.device)]

        self.layers = []                                 # per layer: (Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip), kernel layouts
        for l in range(n):
            w = dilate_weights[l]
            if w.size(2) != 2:
                raise AssertionError("nv-wavenet only supports kernel_size 2 dilated convolutions")
            _expect(f"dilate_weights[{l}]", w, (2 * R, R))
            _expect(f"dilate_biases[{l}]", dilate_biases[l], (2 * R,))
            _expect(f"res_weights[{l}]", res_weights[l], (R, R))
            _expect(f"res_biases[{l}]", res_biases[l], (R,))
            _expect(f"skip_weights[{l}]", skip_weights[l], (S, R))
            _expect(f"skip_biases[{l}]", skip_biases[l], (S,))
            self.layers.append((column_major(w[:, :, 0]), column_major(w[:, :, 1]), dilate_biases[l],
                                column_major(res_weights[l]), res_biases[l], column_major(skip_weights[l]), skip_biases[l]))
        self.num_layers = n

    def infer(self, cond_input, implementation):
        """cond_input: channels x batch x num_layers x samples (pytorch/nv_wavenet.py:172-196); returns int32 [batch][samples]."""
        if (cond_input.size(0), cond_input.size(2)) != (2 * self.R, self.num_layers):
            raise AssertionError(f"Inputs are channels x batch x num_layers x samples; got {tuple(cond_input.size())}")
        batch_size, sample_count = cond_input.size(1), cond_input.size(3)
        lh = column_major(cond_input).float()                       # [samples][layers][batch][2R]
        samples = torch.empty((batch_size, sample_count), dtype=torch.int32, device="cuda")
        f32 = lambda t: t.float().contiguous()
        per_layer = [[f32(t) for t in layer] for layer in self.layers]   # keeps the buffers alive across the call
        column = lambda k: (C.c_void_p * self.num_layers)(*[layer[k].data_ptr() for layer in per_layer])
        emb_prev, emb_cur, conv_out, conv_end = f32(self.embedding_prev), f32(self.embedding_curr), f32(self.conv_out), f32(self.conv_end)
        self._lib.wavenet_infer(sample_count, batch_size, emb_prev.data_ptr(), emb_cur.data_ptr(), self.num_layers, self.max_dilation,
                                column(0), column(1), column(2), column(3), column(4), column(5), column(6),
                                conv_out.data_ptr(), conv_end.data_ptr(), int(bool(self.use_embed_tanh)),
                                lh.data_ptr(), implementation, samples.data_ptr())
        return samples
