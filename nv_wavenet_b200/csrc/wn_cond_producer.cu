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
// fp32 FMA, every output element accumulated in the order of the shared element functions (so host reference, simple kernels
// and the tiled kernel below agree bit for bit).  The projection is 94 % of the flops (C3: 0.42 TFLOP per 16000 samples x 64
// utterances): a register-tiled kernel (128 x 128 outputs per CTA, 8 x 8 per thread, operands through shared memory).
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

// Tiled projection: rows r = j * B + b (sample-major, the order of the output), columns o = l * 2R + c; out = bc[o] + sum_co U[r][co] * Wc[o][co]
// with co ascending (project_element's order).  CTA 256 threads = 16 x 16, thread (ty, tx): rows ty*8 .. +7, columns tx*4 .. +3 and 64 + tx*4 .. +3.
constexpr int PT = 128, PK = 16;
__global__ void __launch_bounds__(256) project_tiled_kernel(float* __restrict__ out, const float* __restrict__ U, const float* __restrict__ Wc, const float* __restrict__ bc,
                                                            int B, int C, int L, int R2, int m)
{
    __shared__ __align__(16) float Us[PK][PT];
    __shared__ __align__(16) float Ws[PK][PT];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long rows = (long)m * B;
    const int cols = L * R2;
    const long r0 = (long)blockIdx.x * PT;
    const int o0 = blockIdx.y * PT;
    // loader: thread -> (tile row lr, 8 consecutive k)
    const int lr = tid >> 1, lk = (tid & 1) * 8;
    const long ur = r0 + lr;
    const float* usrc = nullptr;
    if (ur < rows) { const long j = ur / B; const int b = (int)(ur - j * B); usrc = U + ((size_t)b * m + j) * C; }
    const int wo = o0 + lr;
    const float* wsrc = wo < cols ? Wc + (size_t)wo * C : nullptr;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
#pragma unroll
        for (int q = 0; q < 8; q++) { const int o = o0 + (q < 4 ? tx * 4 + q : 64 + tx * 4 + q - 4); acc[i][q] = o < cols ? bc[o] : 0.f; }
    }
    for (int k0 = 0; k0 < C; k0 += PK) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int k = k0 + lk + q;
            Us[lk + q][lr] = (usrc && k < C) ? usrc[k] : 0.f;
            Ws[lk + q][lr] = (wsrc && k < C) ? wsrc[k] : 0.f;
        }
        __syncthreads();
        const int kn = C - k0 < PK ? C - k0 : PK;
        for (int k = 0; k < kn; k++) {
            const float4 ua = *reinterpret_cast<const float4*>(&Us[k][ty * 8]), ub = *reinterpret_cast<const float4*>(&Us[k][ty * 8 + 4]);
            const float4 wa = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]), wb = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
            const float u[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int q = 0; q < 8; q++) acc[i][q] = fmaf(u[i], w[q], acc[i][q]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const long r = r0 + ty * 8 + i;
        if (r >= rows) continue;
        const long j = r / B; const int b = (int)(r - j * B);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int o = o0 + h * 64 + tx * 4;                // four consecutive columns: same layer (4 | 2R)
            if (o >= cols) continue;
            const int l = o / R2, c = o - l * R2;
            *reinterpret_cast<float4*>(out + (((size_t)j * L + l) * B + b) * R2 + c) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
        }
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
    const int R2 = 2 * R;
    if (R2 % 4 == 0 && (((size_t)out) & 15) == 0) {
        const dim3 grid((unsigned)(((size_t)m * B + PT - 1) / PT), (unsigned)((L * R2 + PT - 1) / PT));
        project_tiled_kernel<<<grid, 256, 0, stream>>>(out, U, Wc, bc, B, C, L, R2, m);
    } else {
        project_kernel<<<grid_for((size_t)m * L * B * R2), 256, 0, stream>>>(out, U, Wc, bc, B, C, L, R2, m);
    }
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
