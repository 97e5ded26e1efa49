// wn_stream_kernel.cu -- "stream" kernel family: one persistent CTA per batch tile runs the WHOLE
// autoregressive loop (embed -> L x (dilated 2x1 conv + gated tanh*sigmoid + 1x1 residual + 1x1 skip
// accumulate) -> two output layers -> softmax -> categorical sample -> feed back) for `count`
// consecutive samples.  Weights are streamed from L2 every step with coalesced column-major reads;
// activations never leave shared memory; the only HBM streams are the conditioning Lh (prefetched one
// layer ahead into registers) and the dilation history ring.
//
// Replaces nv_wavenet_{singleblock,dualblock,persistent}.cuh + matrix_math.cuh + softmax.cuh of the
// reference for
//   * fp32 (TD=float): the BIT-EXACT path.  Operation order is the reference CPU model's
//     (nv_wavenet_reference.cpp:59-121, matrix.cpp:85-183): left-to-right dot products with separately
//     rounded multiply and add, ((a_prev + a_cur) + Bh) + Lh, (Wres.h + Bres) + x, (Wskip.h + skip) + Bskip,
//     softmax with max initialised to 0, sequential sum and p = e / sum, first index with sel < cumsum(p).
//     exp/tanh are the portable double-precision forms of wn_math.cuh.  A CPU evaluation of the same
//     formulas (oracle/wavenet_oracle.c, PORTABLE mode) gives identical bits.
//   * fp16 storage (TD=__half): weights / Lh / ring in fp16, GEMM inputs rounded to fp16, fp32 FMA
//     accumulate, fast MUFU transcendentals.  Fallback for shapes the tensor-core kernel does not cover.
#include "wn_common.h"
#include "wn_math.cuh"

#include <type_traits>

namespace {

template <typename TD> struct Num;
template <> struct Num<float> {
    static constexpr bool exact = true;
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float q(float v) { return v; }
    static __device__ __forceinline__ float mac(float acc, float w, float x) { return __fadd_rn(acc, __fmul_rn(w, x)); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float tanh_(float v) { return wn::tanhf_portable(v); }
    static __device__ __forceinline__ float sigmoid_(float v) { return wn::sigmoidf_portable(v); }
};
template <> struct Num<__half> {
    static constexpr bool exact = false;
    static __device__ __forceinline__ float ld(const __half* p) { return __half2float(__ldg(p)); }
    static __device__ __forceinline__ float ldcg(const __half* p) { return __half2float(__ldcg(p)); }
    static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half_rn(v); }
    static __device__ __forceinline__ float q(float v) { return __half2float(__float2half_rn(v)); }
    static __device__ __forceinline__ float mac(float acc, float w, float x) { return fmaf(w, x, acc); }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float tanh_(float v) { return wn::tanhf_fast(v); }
    static __device__ __forceinline__ float sigmoid_(float v) { return wn::sigmoidf_fast(v); }
};

// fp32 in the REFERENCE GPU KERNELS' arithmetic (NVWN_FP32_FAST): fused multiply-add, two interleaved partial sums per dot
// product (GEMM<R,2>, matrix_math.cuh:80-117), single-precision libm tanh / exp.  Not bit-identical to the CPU model -- it
// agrees with it like the reference's own kernels do (sampled indices equal unless a selector falls within rounding of a
// class boundary; nv_wavenet_test.cu:273-298 tolerances on the activations) -- but free of the 8-cycle-per-term serial chain.
struct NumFast32 {
    static constexpr bool exact = false;
    static __device__ __forceinline__ float ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
    static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
    static __device__ __forceinline__ float q(float v) { return v; }
    static __device__ __forceinline__ float mac(float acc, float w, float x) { return fmaf(w, x, acc); }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
    static __device__ __forceinline__ float tanh_(float v) { return tanhf(v); }
    static __device__ __forceinline__ float sigmoid_(float v) { return 1.f / (1.f + expf(-v)); }
};
template <typename TD, bool FAST> struct NumSel { using type = Num<TD>; };
template <> struct NumSel<float, true> { using type = NumFast32; };

// acc[b] = sum_k W[row + k*M] * xs[b][k], k ascending (matrix.cpp:85-102 order).
// KB weights are requested back to back before the first use, so the L2 latency is paid once per
// KB columns instead of once per 8.
template <typename TD, int BT, int KB, bool FAST>
__device__ __forceinline__ void dot_cols(const TD* __restrict__ W, int M, int K, int row,
                                         const float* __restrict__ xs, float (&acc)[BT])
{
    using N = typename NumSel<TD, FAST>::type;
    float odd[BT];                                   // second partial sum of the non-exact contracts (odd k)
#pragma unroll
    for (int b = 0; b < BT; b++) { acc[b] = 0.f; odd[b] = 0.f; }
    const TD* wp = W + row;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += KB) {
        float w[KB];
#pragma unroll
        for (int j = 0; j < KB; j++) w[j] = N::ld(wp + (size_t)(k0 + j) * M);
#pragma unroll
        for (int j = 0; j < KB; j += 4) {
#pragma unroll
            for (int b = 0; b < BT; b++) {
                const float4 xa = *reinterpret_cast<const float4*>(xs + b * K + k0 + j);
                if (N::exact) {
                    float a = acc[b];
                    a = N::mac(a, w[j], xa.x); a = N::mac(a, w[j + 1], xa.y); a = N::mac(a, w[j + 2], xa.z); a = N::mac(a, w[j + 3], xa.w);
                    acc[b] = a;
                } else {
                    float a = acc[b], o = odd[b];
                    a = N::mac(a, w[j], xa.x); o = N::mac(o, w[j + 1], xa.y); a = N::mac(a, w[j + 2], xa.z); o = N::mac(o, w[j + 3], xa.w);
                    acc[b] = a; odd[b] = o;
                }
            }
        }
    }
    if (!N::exact) {
#pragma unroll
        for (int b = 0; b < BT; b++) acc[b] += odd[b];
    }
}

constexpr int cmax(int a, int b) { return a > b ? a : b; }

template <int R, int S> struct Shape {
    static constexpr int NT = cmax(cmax(4 * R, R + S), 128);       // threads per CTA
};

template <int R, int S, int BT>
__host__ __device__ constexpr size_t stream_smem_floats(int A, int L)
{
    return (size_t)BT * (3 * R + 4 * R + R + 2 * S + 3 * A) + BT * 4 + BT * 2 + L + 4;
}

template <typename TD, int R, int S, int BT, bool FAST>
__global__ void __launch_bounds__(Shape<R, S>::NT, 1) wn_stream_kernel(const WnParams p)
{
    using N = typename NumSel<TD, FAST>::type;
    constexpr int NT = Shape<R, S>::NT;
    constexpr int NACT = 2 * R * BT;
    constexpr int ACT_PER = (NACT + NT - 1) / NT;
    constexpr int KBR = R < 64 ? R : 64;          // weight columns in flight per thread in the R-deep dots
    static_assert(R * BT <= NT, "one x-task per thread");

    const int tid = threadIdx.x;
    const int A = p.A, L = p.L, B = p.B;
    const int b0 = blockIdx.x * BT;
    const int slots = p.maxDil + 1;

    extern __shared__ __align__(16) float sm[];
    float* x = sm;                         // [BT][R]   residual stream, fp32
    float* xq = x + BT * R;                // [BT][R]   GEMM input of the current layer
    float* xp = xq + BT * R;               // [BT][R]   GEMM input x[t-d]
    float* ap = xp + BT * R;               // [BT][2R]  Wprev . x[t-d]
    float* ac = ap + BT * 2 * R;           // [BT][2R]  Wcur . x[t]   -> overwritten by tanh / sigmoid values
    float* hq = ac + BT * 2 * R;           // [BT][R]   gated activation, GEMM input
    float* skip = hq + BT * R;             // [BT][S]   running skip sum, fp32
    float* skq = skip + BT * S;            // [BT][S]   relu(skip) as GEMM input
    float* zsq = skq + BT * S;             // [BT][A]
    float* za = zsq + BT * A;              // [BT][A]   logits
    float* ex = za + BT * A;               // [BT][A]   exp / p
    float* red = ex + BT * A;              // [BT][4]   max, sum
    int* ysm = reinterpret_cast<int*>(red + BT * 4);   // [BT][2]  yPrev, yCur
    int* dil = ysm + BT * 2;               // [L]

    const TD* embPrev = static_cast<const TD*>(p.embPrev);
    const TD* embCur = static_cast<const TD*>(p.embCur);
    const TD* Wprev = static_cast<const TD*>(p.Wprev);
    const TD* Wcur = static_cast<const TD*>(p.Wcur);
    const TD* Wres = static_cast<const TD*>(p.Wres);
    const TD* Wskip = static_cast<const TD*>(p.Wskip);
    const TD* Wzs = static_cast<const TD*>(p.Wzs);
    const TD* Wza = static_cast<const TD*>(p.Wza);
    const TD* Bh = static_cast<const TD*>(p.Bh);
    const TD* Bres = static_cast<const TD*>(p.Bres);
    const TD* Bskip = static_cast<const TD*>(p.Bskip);
    const TD* Bzs = static_cast<const TD*>(p.Bzs);
    const TD* Bza = static_cast<const TD*>(p.Bza);
    const TD* Lh = static_cast<const TD*>(p.Lh);
    TD* ring = static_cast<TD*>(p.ring);

    // dilation schedule 1,2,4..maxDil,1,2..  (nv_wavenet.cuh:99-111, reference.cpp:285-289)
    if (tid == 0) {
        int d = 1;
        for (int l = 0; l < L; l++) { dil[l] = d; d <<= 1; if (d > p.maxDil) d = 1; }
    }
    if (tid < BT) { ysm[tid * 2] = p.yPrev[b0 + tid]; ysm[tid * 2 + 1] = p.yCur[b0 + tid]; }
    __syncthreads();

    // fixed per-thread roles
    const bool has_x = tid < R * BT;
    const int xb = has_x ? tid / R : 0, xr = has_x ? tid % R : 0;

    const int t_end = p.init_sample + p.count;
    auto ring_at = [&](int t, int l, int b, int r) -> TD* {
        return ring + (((size_t)(t % slots) * L + l) * B + (b0 + b)) * R + r;
    };
    auto lh_at = [&](int t, int l, int idx) -> const TD* {
        const int b = idx / (2 * R), row = idx % (2 * R);
        return Lh + (((size_t)t * L + l) * B + (b0 + b)) * (2 * R) + row;
    };

    // prefetch registers for (t = init, l = 0)
    float lh_nxt[ACT_PER], bh_nxt[ACT_PER];
    float xp_nxt = 0.f;
    {
        const int t = p.init_sample;
#pragma unroll
        for (int j = 0; j < ACT_PER; j++) {
            const int idx = tid + j * NT;
            lh_nxt[j] = (idx < NACT && t < t_end) ? N::ldcg(lh_at(t, 0, idx)) : 0.f;
            bh_nxt[j] = (idx < NACT) ? N::ld(Bh + idx % (2 * R)) : 0.f;
        }
        if (has_x) xp_nxt = (t >= 1) ? N::ldcg(ring_at(t - 1, 0, xb, xr)) : 0.f;
    }

    for (int t = p.init_sample; t < t_end; t++) {
        const bool dump = p.dump && (t == t_end - 1);
        // ---- embedding (nv_wavenet_reference.cpp:42-57) ----
        if (has_x) {
            const int yp = ysm[xb * 2], yc = ysm[xb * 2 + 1];
            float e = N::add(N::ld(embPrev + (size_t)yp * R + xr), N::ld(embCur + (size_t)yc * R + xr));
            if (p.tanhEmbed) e = N::tanh_(e);
            x[xb * R + xr] = e;
            xq[xb * R + xr] = N::q(e);
            N::st(ring_at(t, 0, xb, xr), e);
            xp[xb * R + xr] = xp_nxt;
        }
        for (int i = tid; i < S * BT; i += NT) skip[i] = 0.f;     // zero matrix (reference.cpp:290)
        __syncthreads();

        for (int l = 0; l < L; l++) {
            // ---- conditioning / history prefetch for the next layer (or layer 0 of the next sample) ----
            float lh_cur[ACT_PER], bh_cur[ACT_PER];
#pragma unroll
            for (int j = 0; j < ACT_PER; j++) { lh_cur[j] = lh_nxt[j]; bh_cur[j] = bh_nxt[j]; }
            {
                const bool wrap = (l + 1 == L);
                const int tn = wrap ? t + 1 : t, ln = wrap ? 0 : l + 1;
                const bool live = tn < t_end;
#pragma unroll
                for (int j = 0; j < ACT_PER; j++) {
                    const int idx = tid + j * NT;
                    lh_nxt[j] = (live && idx < NACT) ? N::ldcg(lh_at(tn, ln, idx)) : 0.f;
                    bh_nxt[j] = (idx < NACT) ? N::ld(Bh + (size_t)ln * 2 * R + idx % (2 * R)) : 0.f;
                }
                // x[t-d] of the next layer.  For (t+1, layer 0) the source is this sample's embedding,
                // written above by this CTA and ordered by the barriers in between.
                const int dn = dil[ln];
                xp_nxt = (live && has_x && tn >= dn) ? N::ldcg(ring_at(tn - dn, ln, xb, xr)) : 0.f;
            }

            // ---- stage 1: a_prev = Wprev.x[t-d], a_cur = Wcur.x[t]  (reference.cpp:61-65) ----
            if (tid < 4 * R) {
                const bool cur = tid >= 2 * R;
                const int row = cur ? tid - 2 * R : tid;
                const TD* W = (cur ? Wcur : Wprev) + (size_t)l * 2 * R * R;
                float acc[BT];
                dot_cols<TD, BT, KBR, FAST>(W, 2 * R, R, row, cur ? xq : xp, acc);
                float* dst = cur ? ac : ap;
#pragma unroll
                for (int b = 0; b < BT; b++) dst[b * 2 * R + row] = acc[b];
            }
            __syncthreads();

            // ---- pre-activation adds and tanh / sigmoid (reference.cpp:67-72, 76-78) ----
#pragma unroll
            for (int j = 0; j < ACT_PER; j++) {
                const int idx = tid + j * NT;
                if (idx < NACT) {
                    const int row = idx % (2 * R);
                    float v = N::add(ap[idx], ac[idx]);
                    v = N::add(v, bh_cur[j]);
                    v = N::add(v, lh_cur[j]);
                    ac[idx] = (row < R) ? N::tanh_(v) : N::sigmoid_(v);
                }
            }
            __syncthreads();

            // ---- h = tanh * sigmoid; stage-1 inputs are dead, install x[t-d] of the next layer ----
            if (has_x) {
                const float h = N::mul(ac[xb * 2 * R + xr], ac[xb * 2 * R + xr + R]);
                hq[xb * R + xr] = N::q(h);
                xp[xb * R + xr] = xp_nxt;
            }
            __syncthreads();

            // ---- stage 3: residual (reference.cpp:82-84) and skip (reference.cpp:86-90) ----
            if (tid < R + S) {
                const bool is_res = tid < R;
                const int row = is_res ? tid : tid - R;
                const TD* W = is_res ? Wres + (size_t)l * R * R : Wskip + (size_t)l * S * R;
                const float bias = is_res ? N::ld(Bres + (size_t)l * R + row) : N::ld(Bskip + (size_t)l * S + row);
                float acc[BT];
                dot_cols<TD, BT, KBR, FAST>(W, is_res ? R : S, R, row, hq, acc);
                if (is_res) {
#pragma unroll
                    for (int b = 0; b < BT; b++) {
                        float v = N::add(acc[b], bias);
                        v = N::add(v, x[b * R + row]);
                        x[b * R + row] = v;
                        xq[b * R + row] = N::q(v);
                        if (l + 1 < L) N::st(ring_at(t, l + 1, b, row), v);
                        if (dump) p.xtOut[((size_t)l * B + b0 + b) * R + row] = v;
                    }
                } else {
#pragma unroll
                    for (int b = 0; b < BT; b++) {
                        float v = N::add(acc[b], skip[b * S + row]);
                        v = N::add(v, bias);
                        if (l == L - 1) { v = (v < 0.f) ? 0.f : v; skq[b * S + row] = N::q(v); }
                        skip[b * S + row] = v;
                        if (dump) p.skipOut[((size_t)l * B + b0 + b) * S + row] = v;
                    }
                }
            }
            __syncthreads();
        }

        // ---- output layers (reference.cpp:93-104) ----
        for (int row = tid; row < A; row += NT) {
            float acc[BT];
            const float bias = N::ld(Bzs + row);
            dot_cols<TD, BT, 32, FAST>(Wzs, A, S, row, skq, acc);
#pragma unroll
            for (int b = 0; b < BT; b++) {
                float v = N::add(acc[b], bias);
                v = (v < 0.f) ? 0.f : v;
                zsq[b * A + row] = N::q(v);
                if (dump) p.Zs[(size_t)(b0 + b) * A + row] = v;
            }
        }
        __syncthreads();
        for (int row = tid; row < A; row += NT) {
            float acc[BT];
            const float bias = N::ld(Bza + row);
            dot_cols<TD, BT, 32, FAST>(Wza, A, A, row, zsq, acc);
#pragma unroll
            for (int b = 0; b < BT; b++) {
                const float v = N::add(acc[b], bias);
                za[b * A + row] = v;
                if (dump) p.Za[(size_t)(b0 + b) * A + row] = v;
            }
        }
        __syncthreads();

        // ---- softmax + categorical sample ----
        const int warp = tid >> 5, lane = tid & 31;
        if (N::exact) {
            // matrix.cpp:167-183 and reference.cpp:106-121, bit for bit
            if (warp < BT) {
                float mx = 0.f;                                      // "float max = 0.f"
                for (int a = lane; a < A; a += 32) mx = fmaxf(mx, za[warp * A + a]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if (lane == 0) red[warp * 4] = mx;
            }
            __syncthreads();
            for (int i = tid; i < A * BT; i += NT) ex[i] = wn::expf_portable(__fsub_rn(za[i], red[(i / A) * 4]));
            __syncthreads();
            if (tid < BT) {
                float s = 0.f;
                const float* e = ex + tid * A;
                for (int a = 0; a < A; a += 4) {
                    const float4 v = *reinterpret_cast<const float4*>(e + a);
                    s = __fadd_rn(s, v.x); s = __fadd_rn(s, v.y); s = __fadd_rn(s, v.z); s = __fadd_rn(s, v.w);
                }
                red[tid * 4 + 1] = s;
            }
            __syncthreads();
            for (int i = tid; i < A * BT; i += NT) {
                const float pr = __fdiv_rn(ex[i], red[(i / A) * 4 + 1]);
                ex[i] = pr;
                if (dump) p.P[(size_t)(b0 + i / A) * A + (i % A)] = pr;
            }
            __syncthreads();
            if (tid < BT) {
                const float sel = p.sel[(size_t)t * B + b0 + tid];
                const float* pr = ex + tid * A;
                float cs = 0.f;
                int y = -1;
                for (int a = 0; a < A; a++) {
                    cs = __fadd_rn(cs, pr[a]);
                    if (sel < cs) { y = a; break; }
                }
                if (y < 0) y = A - 1;       // the reference asserts here (reference.cpp:119)
                p.yOut[(size_t)(b0 + tid) * p.N + t] = y;
                const int fb = p.forced ? p.forced[(size_t)(b0 + tid) * p.N + t] : y;
                ysm[tid * 2] = ysm[tid * 2 + 1];
                ysm[tid * 2 + 1] = fb;
            }
        } else {
            // one warp per utterance; lane owns A/32 consecutive rows so that the scan is in row order
            if (warp < BT) {
                const int per = A / 32;
                const float* z = za + warp * A + lane * per;
                float mx = 0.f;
                for (int j = 0; j < per; j++) mx = fmaxf(mx, z[j]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                float* e = ex + warp * A + lane * per;
                float ls = 0.f;
                for (int j = 0; j < per; j++) {
                    const float v = wn::exp2f_fast((z[j] - mx) * 1.4426950408889634f);
                    e[j] = v;
                    ls += v;
                }
                float inc = ls;                                     // inclusive scan over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const float total = __shfl_sync(0xffffffffu, inc, 31);
                const float target = p.sel[(size_t)t * B + b0 + warp] * total;
                const unsigned hit = __ballot_sync(0xffffffffu, target < inc);
                int y = A - 1;
                if (hit) {
                    const int first = __ffs(hit) - 1;
                    if (lane == first) {
                        float cs = inc - ls;
                        y = lane * per + per - 1;
                        for (int j = 0; j < per; j++) {
                            cs += e[j];
                            if (target < cs) { y = lane * per + j; break; }
                        }
                    }
                    y = __shfl_sync(0xffffffffu, y, first);
                }
                if (dump) {
                    const float inv = 1.f / total;
                    for (int j = 0; j < per; j++) p.P[(size_t)(b0 + warp) * A + lane * per + j] = e[j] * inv;
                }
                if (lane == 0) {
                    p.yOut[(size_t)(b0 + warp) * p.N + t] = y;
                    const int fb = p.forced ? p.forced[(size_t)(b0 + warp) * p.N + t] : y;
                    ysm[warp * 2] = ysm[warp * 2 + 1];
                    ysm[warp * 2 + 1] = fb;
                }
            }
        }
        __syncthreads();
    }

    if (tid < BT) { p.yPrev[b0 + tid] = ysm[tid * 2]; p.yCur[b0 + tid] = ysm[tid * 2 + 1]; }
}

template <typename TD, int R, int S, int BT, bool FAST>
cudaError_t launch_one(const WnParams& p, cudaStream_t stream, WnLaunchInfo* info)
{
    constexpr int NT = Shape<R, S>::NT;
    const size_t smem = stream_smem_floats<R, S, BT>(p.A, p.L) * sizeof(float);
    auto kfn = wn_stream_kernel<TD, R, S, BT, FAST>;
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int grid = p.B / BT;
    kfn<<<grid, NT, smem, stream>>>(p);
    if (info) { info->grid = grid; info->block = NT; info->smem_bytes = (int)smem; info->batch_per_cta = BT; info->cluster = 1; }
    return cudaGetLastError();
}

template <typename TD, int R, int S, bool FAST>
cudaError_t launch_bt(const WnParams& p, int bt, cudaStream_t stream, WnLaunchInfo* info)
{
    switch (bt) {
        case 4: return launch_one<TD, R, S, 4, FAST>(p, stream, info);
        case 2: return launch_one<TD, R, S, 2, FAST>(p, stream, info);
        default: return launch_one<TD, R, S, 1, FAST>(p, stream, info);
    }
}

template <typename TD, bool FAST>
cudaError_t launch_shape(const WnParams& p, int bt, cudaStream_t stream, WnLaunchInfo* info)
{
    if (p.R == 32 && p.S == 128) return launch_bt<TD, 32, 128, FAST>(p, bt, stream, info);
    if (p.R == 64 && p.S == 128) return launch_bt<TD, 64, 128, FAST>(p, bt, stream, info);
    if (p.R == 64 && p.S == 256) return launch_bt<TD, 64, 256, FAST>(p, bt, stream, info);
    if (p.R == 128 && p.S == 256) return launch_bt<TD, 128, 256, FAST>(p, bt, stream, info);
    return cudaErrorInvalidValue;
}

}  // namespace

bool wn_stream_supported(int R, int S, int A, bool)
{
    const bool shape = (R == 32 && S == 128) || (R == 64 && S == 128) || (R == 64 && S == 256) || (R == 128 && S == 256);
    return shape && A % 32 == 0 && A >= 32 && A <= 4096;
}

// Batch tile per CTA: every CTA re-reads all weights from L2 each sample, so L2 traffic per step is
// (B / BT) * weight_bytes; keep that under ~128 MB/step while using as many SMs as possible.
static int pick_bt(const WnParams& p, bool fp16)
{
    if (const char* env = getenv("NVWN_STREAM_BT")) {
        const int v = atoi(env);
        if ((v == 1 || v == 2 || v == 4) && p.B % v == 0) return v;
    }
    const double wbytes = (fp16 ? 2.0 : 4.0) * ((double)p.L * (5.0 * p.R * p.R + (double)p.S * p.R) + (double)p.A * p.S + (double)p.A * p.A);
    int bt = 1;
    while (bt < 4 && p.B % (bt * 2) == 0 && ((double)(p.B / bt) * wbytes > 128e6 || p.B / bt > 296)) bt *= 2;
    return bt;
}

// contract: 0 = fp32 bit-exact (CPU model's operation order), 1 = fp16 storage, 2 = fp32 in the reference GPU kernels' arithmetic
cudaError_t wn_launch_stream(const WnParams& p, int contract, cudaStream_t stream, WnLaunchInfo* info)
{
    const bool fp16 = contract == 1;
    if (!wn_stream_supported(p.R, p.S, p.A, fp16)) return cudaErrorInvalidValue;
    const int bt = pick_bt(p, fp16);
    if (info) info->kernel = 16;
    if (fp16) return launch_shape<__half, false>(p, bt, stream, info);
    return contract == 2 ? launch_shape<float, true>(p, bt, stream, info) : launch_shape<float, false>(p, bt, stream, info);
}
