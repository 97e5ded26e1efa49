"""CPU part of the pytorch/nv_wavenet.py mirror (nv_wavenet_b200/nv_wavenet.py): layouts prepared by the constructor
and the shape checks.  (The GPU part -- infer() against the oracle -- is in tests/test_gpu_parity.py.)"""
import numpy as np
import pytest
import torch

from nv_wavenet_b200.nv_wavenet import Impl, NVWaveNet, column_major


def test_column_major_layouts():
    m = torch.arange(6.).reshape(2, 3)
    assert column_major(m).flatten().tolist() == [0, 3, 1, 4, 2, 5]           # column-major order of a [2][3] matrix
    assert torch.equal(column_major(m[:, :, None]), column_major(m))          # conv weight with kernel size 1
    v = torch.arange(4.)
    assert column_major(v) is v
    c = torch.arange(24.).reshape(2, 3, 4, 1)
    assert column_major(c).shape == (1, 4, 3, 2) and column_major(c)[0, 3, 2, 1] == c[1, 2, 3, 0]
    with pytest.raises(AssertionError):
        column_major(torch.zeros(2, 2, 2))


def _weights(L, R, S, A, rng):
    t = lambda *s: torch.from_numpy(rng.standard_normal(s).astype(np.float32))
    return dict(embedding_prev=t(A, R), embedding_curr=t(A, R), conv_out_weight=t(A, S, 1), conv_end_weight=t(A, A, 1),
                dilate_weights=[t(2 * R, R, 2) for _ in range(L)], dilate_biases=[t(2 * R) for _ in range(L)], max_dilation=4,
                res_weights=[t(R, R, 1) for _ in range(L - 1)], res_biases=[t(R) for _ in range(L - 1)],
                skip_weights=[t(S, R, 1) for _ in range(L)], skip_biases=[t(S) for _ in range(L)], use_embed_tanh=False)


def test_constructor_prepares_kernel_layouts():
    from nv_wavenet_b200 import _lib
    lib = _lib.lib()
    R, S, A = lib.get_R(), lib.get_S(), lib.get_A()
    L = 3
    w = _weights(L, R, S, A, np.random.default_rng(0))
    net = NVWaveNet(**w)
    assert (net.R, net.S, net.A, net.num_layers) == (R, S, A, L) and Impl.PERSISTENT == 3
    # embeddings are consumed as emb[a * R + r]
    assert torch.equal(net.embedding_prev.flatten(), w["embedding_prev"].flatten())
    # matrices column-major: element (row, k) at k * M + row
    wp, wc, bh, wr, br, ws, bs = net.layers[1]
    d = w["dilate_weights"][1]
    assert wp.flatten()[5 * 2 * R + 7] == d[7, 5, 0] and wc.flatten()[5 * 2 * R + 7] == d[7, 5, 1]
    assert wr.flatten()[3 * R + 9] == w["res_weights"][1][9, 3, 0] and ws.flatten()[2 * S + 11] == w["skip_weights"][1][11, 2, 0]
    assert torch.equal(bh, w["dilate_biases"][1]) and torch.equal(bs, w["skip_biases"][1])
    assert net.conv_out.flatten()[4 * A + 6] == w["conv_out_weight"][6, 4, 0]
    # the appended residual layer of the last layer is all zero
    assert torch.count_nonzero(net.layers[L - 1][3]) == 0 and torch.count_nonzero(net.layers[L - 1][4]) == 0


def test_constructor_rejects_wrong_shapes():
    from nv_wavenet_b200 import _lib
    lib = _lib.lib()
    R, S, A = lib.get_R(), lib.get_S(), lib.get_A()
    rng = np.random.default_rng(1)
    w = _weights(2, R, S, A, rng); w["embedding_prev"] = torch.zeros(A, R + 1)
    with pytest.raises(AssertionError):
        NVWaveNet(**w)
    w = _weights(2, R, S, A, rng); w["skip_biases"] = w["skip_biases"][:1]
    with pytest.raises(AssertionError):
        NVWaveNet(**w)
    w = _weights(2, R, S, A, rng); w["dilate_weights"][0] = torch.zeros(2 * R, R, 3)
    with pytest.raises(AssertionError):
        NVWaveNet(**w)
