// wn_cond_producer.cu -- conditioning producer: local-conditioning features (mel frames) ->
// Lh[n][l][b][2R] on the device, in chunks of whole samples that go straight into the engine's conditioning store
// (nvwn_set_conditioning), so the [N][L][B][2R] fp32 tensor of the reference (10.5 GB for C3) never exists.
//
// Replaces, for inference, WaveNet.get_cond_input (pytorch/wavenet.py:190-202) and the permutes that follow it
// (pytorch/nv_wavenet.py:48-49,181):
//     u = ConvTranspose1d(C, C, window, stride)(features)[:, :, :-(window - stride)]      -> [B][C][T * stride]
//     y = Conv1d(C, L * 2R, 1)(u)                                                       -> [B][L * 2R][T * stride]
//     Lh[n][l][b][c] = y[b][l * 2R + c][n]
// Two stages (the composed operator would need C * window * L * 2R weights).  The arithmetic of one output element
// lives in __host__ __device__ functions that the host reference below (nvwn_cond_from_features_host, used by the CPU
// tests against vectors generated from the reference's own module) and the kernels share.
// Plain fp32 FMA loops: both stages together are ~25 % of the model's flops but embarrassingly parallel and one-off
// per utterance batch; they are bound by the fp32 store of the chunk, not worth tensor cores.
#include "wn_common.h"

namespace {

// u[b][co][n] of the trimmed transposed convolution.  features [B][C][T], Wu [C_in][C_out][K] (torch ConvTranspose1d
// layout) or, with WU_T, its [K][C_in][C_out] transpose (coalesced across co).  out[n] += in[m] * W[k] for m*stride + k = n.
template <bool WU_T>
__host__ __device__ inline float upsample_element(const float* __restrict__ feat, const float* __restrict__ Wu, const float* __restrict__ bu,
                                                  int C, int T, int K, int stride, int b, int co, int n)
{
    float acc = bu[co];
    for (int m = n / stride; m >= 0; m--) {
        const int k = n - m * stride;
        if (k >= K) break;
        if (m >= T) continue;
        for (int ci = 0; ci < C; ci++) {
            const float w = WU_T ? Wu[((size_t)k * C + ci) * C + co] : Wu[((size_t)ci * C + co) * K + k];
            acc = fmaf(feat[((size_t)b * C + ci) * T + m], w, acc);
        }
    }
    return acc;
}

// y[b][o][n] of the 1x1 convolution from u stored as U[b][n][C] (channel fastest); Wc [L*2R][C]
__host__ __device__ inline float project_element(const float* __restrict__ Urow, const float* __restrict__ Wc, const float* __restrict__ bc, int C, int o)
{
    float acc = bc[o];
    const float* w = Wc + (size_t)o * C;
    for (int co = 0; co < C; co++) acc = fmaf(Urow[co], w[co], acc);
    return acc;
}

__global__ void transpose_wu_kernel(float* __restrict__ dst, const float* __restrict__ src, int C, int K)
{
    const size_t total = (size_t)C * C * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % C), ci = (int)((i / C) % C), k = (int)(i / ((size_t)C * C));     // dst index [k][ci][co]
        dst[i] = src[((size_t)ci * C + co) * K + k];
    }
}

// U[b][j][co] for samples n0 + j, j < m
__global__ void upsample_kernel(float* __restrict__ U, const float* __restrict__ feat, const float* __restrict__ WuT, const float* __restrict__ bu,
                                int B, int C, int T, int K, int stride, int n0, int m)
{
    const size_t total = (size_t)B * m * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(i % C), j = (int)((i / C) % m), b = (int)(i / ((size_t)C * m));
        U[i] = upsample_element<true>(feat, WuT, bu, C, T, K, stride, b, co, n0 + j);
    }
}

// out[j][l][b][c] (the engine's fp32 conditioning layout for samples n0 + j) from U[b][j][C]
__global__ void project_kernel(float* __restrict__ out, const float* __restrict__ U, const float* __restrict__ Wc, const float* __restrict__ bc,
                               int B, int C, int L, int R2, int m)
{
    const size_t total = (size_t)m * L * B * R2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % R2), b = (int)((i / R2) % B), l = (int)((i / ((size_t)R2 * B)) % L), j = (int)(i / ((size_t)R2 * B * L));
        out[i] = project_element(U + ((size_t)b * m + j) * C, Wc, bc, C, l * R2 + c);
    }
}

unsigned grid_for(size_t total)
{
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    return (unsigned)(blocks ? blocks : 1);
}

}  // namespace

// device pointers throughout; WuT = [K][C][C] transpose made by wn_cond_transpose_wu; U scratch holds B * m * C floats
cudaError_t wn_cond_transpose_wu(float* WuT, const float* Wu, int C, int K, cudaStream_t stream)
{
    transpose_wu_kernel<<<grid_for((size_t)C * C * K), 256, 0, stream>>>(WuT, Wu, C, K);
    return cudaGetLastError();
}

cudaError_t wn_cond_produce(float* out, float* U, const float* feat, const float* WuT, const float* bu, const float* Wc, const float* bc,
                            int B, int C, int T, int K, int stride, int L, int R, int n0, int m, cudaStream_t stream)
{
    if (m <= 0) return cudaSuccess;
    upsample_kernel<<<grid_for((size_t)B * m * C), 256, 0, stream>>>(U, feat, WuT, bu, B, C, T, K, stride, n0, m);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    project_kernel<<<grid_for((size_t)m * L * B * 2 * R), 256, 0, stream>>>(out, U, Wc, bc, B, C, L, 2 * R, m);
    return cudaGetLastError();
}

// Host reference of the same arithmetic (no GPU): Lh [T*stride][L][B][2R]
void wn_cond_host(float* Lh, const float* feat, const float* Wu, const float* bu, const float* Wc, const float* bc,
                  int B, int C, int T, int K, int stride, int L, int R)
{
    const int Nn = T * stride, R2 = 2 * R;
    float* Urow = new float[C];
    for (int b = 0; b < B; b++)
        for (int n = 0; n < Nn; n++) {
            for (int co = 0; co < C; co++) Urow[co] = upsample_element<false>(feat, Wu, bu, C, T, K, stride, b, co, n);
            for (int l = 0; l < L; l++)
                for (int c = 0; c < R2; c++)
                    Lh[(((size_t)n * L + l) * B + b) * R2 + c] = project_element(Urow, Wc, bc, C, l * R2 + c);
        }
    delete[] Urow;
}
