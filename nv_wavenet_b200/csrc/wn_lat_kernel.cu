// wn_lat_kernel.cu -- latency-mode fp16 kernel of the WaveNet inference loop (sm_100a).
//
// The autoregressive loop is a chain of ~45 tiny dependent GEMMs per sample (M = utterances, N <= 256, K <= 256).  On such a
// chain the tcgen05 round trip (MMA -> commit -> mbarrier -> tcgen05.ld, ~500 cycles) is the cost, not the tensor pipe
// (wn_tc_kernel.cu: 2.4 k cycles per layer).  This kernel keeps the whole chain in REGISTERS: warp-level mma.sync
// (m16n8k16, fp16 x fp16 -> fp32) with the utterances as the M dimension, so that the accumulator fragment of one GEMM is
// -- after the row-local epilogue -- exactly the A fragment of the next one; the only exchange between the eight compute
// warps is one 2 KB shared-memory tile + one named barrier per GEMM stage.
//
//   one persistent CTA per tile of 16 utterances (B = 64 -> 4 SMs), 8 compute warps + 1 producer warp:
//   compute warp w  owns output channels [8w, 8w+8) (+R for the sigmoid half) of every layer GEMM, [32w, 32w+32) of the skip
//                   sum / Zs / Za; residual stream (fp32) and skip sum (fp32) never leave its registers.
//                   per layer:  a = Wcur.x + Wprev.x[t-d] + (Bh + Lh)  ->  h = tanh * sigmoid  -> [h tile, barrier]
//                               x' = Wres.h + Bres + x -> [x tile, history ring, barrier];  skip += Wskip.h
//                   per sample: relu(skip) -> Zs -> Za -> softmax + categorical sample (warp-local, two utterances per warp)
//                               -> embedding gather from the shared-memory resident table.
//   producer warp   streams the weight image (pre-arranged in mma B-fragment order, so that every weight load of a warp is one
//                   conflict-free 512-byte LDS.128) from L2 through a 2-stage shared-memory ring with bulk TMA + mbarriers,
//                   one stage per layer, the four matrices of a layer on four barriers.
//   Lh (the only HBM stream), biases and the dilated history x[t-d] are prefetched into registers one / two layers ahead.
//
// Replaces nv_wavenet_{singleblock,dualblock,persistent}.cuh + matrix_math.cuh + softmax.cuh of the reference for
// T_data = half while the batch is small enough to be latency-bound.  Numerical contract: oracle/wavenet_oracle.c
// WNO_PREC_FP16 (GEMM inputs fp16, fp32 accumulation, fp32 residual stream / skip sum / softmax), gate evaluated with
// tanh.approx.f16x2; no weight folding.
#include "wn_common.h"
#include "wn_math.cuh"
#include "wn_sm100.cuh"

#include <stdlib.h>

namespace {

using namespace sm100;

constexpr int R = 64, A = 256;
constexpr int NCW = 8;                      // compute warps
constexpr int NCT = NCW * 32;
constexpr int NT = NCT + 32;                // + producer warp
constexpr int TU = 16;                      // utterances per tile = M of mma.m16n8k16
constexpr int LROW = 264;                   // padded row (floats) of the transposed-logits buffer
constexpr int EROW = 33;                    // padded row (32-bit words) of the shared-memory embedding table
constexpr int MAXL = 64;

template <int S>
struct Cfg {
    static constexpr int W_PREV = 0, W_CUR = 16384, W_RES = 32768, W_SKIP = 40960;
    static constexpr int LAYER_BYTES = W_SKIP + S * 128;            // = one ring stage
    static constexpr int OJP = S == 256 ? 4 : 2;                     // k-step pairs per output-GEMM stage load
    static constexpr int OLOAD = 32 * OJP * 512;                     // bytes per output-GEMM stage load (all 32 n-tiles)
    static constexpr int NQ_ZS = (S / 32) / OJP, NQ_ZA = (A / 32) / OJP;
    static constexpr int NSK = S / 64;                               // skip n-tiles per warp
    // shared memory map
    static constexpr uint32_t O_RING = 0;
    static constexpr uint32_t O_EMB = 2 * LAYER_BYTES;
    static constexpr uint32_t O_BOUT = O_EMB + A * EROW * 4;         // fp32: Bskip total [S], Bzs [A], Bza [A]
    static constexpr uint32_t O_XBUF = O_BOUT + (S + 2 * A) * 4;
    static constexpr uint32_t O_HBUF = O_XBUF + 2048;
    static constexpr uint32_t O_EPBUF = O_HBUF + 2048;
    static constexpr uint32_t O_OB0 = O_EPBUF + 2048;
    static constexpr uint32_t O_OB1 = O_OB0 + (S / 16) * 512;
    static constexpr uint32_t O_LBUF = O_OB1 + (A / 16) * 512;
    static constexpr uint32_t O_DIL = O_LBUF + TU * LROW * 4;
    static constexpr uint32_t O_YS = O_DIL + MAXL * 4;
    static constexpr uint32_t O_BAR = O_YS + 2 * TU * 4;
    static constexpr uint32_t SMEM = O_BAR + 16 * 8;
};

struct LatImage {
    size_t layer_bytes, off_zs, off_za, off_bias, total;
    size_t b_layer, b_skpre, b_bzs, b_bza;      // float offsets inside the bias block
};
__host__ __device__ inline LatImage lat_image(int S, int L)
{
    LatImage im;
    im.layer_bytes = 40960 + (size_t)S * 128;
    im.off_zs = (size_t)L * im.layer_bytes;
    im.off_za = im.off_zs + (size_t)A * S * 2;
    im.off_bias = im.off_za + (size_t)A * A * 2;
    im.b_layer = 0;                             // [L][8 warps][4 t][8]: Bh tanh pair, Bh sigmoid pair, Bres pair, 0, 0
    im.b_skpre = (size_t)L * 256;               // [L][S] running sum of the skip biases
    im.b_bzs = im.b_skpre + (size_t)L * S;
    im.b_bza = im.b_bzs + A;
    im.total = im.off_bias + (im.b_bza + A) * sizeof(float);
    return im;
}

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b)
{
    asm volatile("st.shared.v2.u32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void load_a(uint32_t (&a)[4], uint32_t addr)
{
    const uint4 v = lds128(addr);
    a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
}
// global loads of data written earlier by this CTA (history ring): L2 only
__device__ __forceinline__ uint4 ldg_cg_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void stg_v2(void* p, uint32_t a, uint32_t b) { asm volatile("st.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void stg_v4(void* p, uint4 v) { asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }
__device__ __forceinline__ uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 h2(uint32_t v) { return *reinterpret_cast<__half2*>(&v); }

// mbarrier by shared-memory address
__device__ __forceinline__ void mbar_init_a(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_a(uint32_t bar, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_a(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
static __device__ __noinline__ void lat_timeout(uint32_t bar, uint32_t parity)
{
    printf("wn_lat: mbarrier wait timed out: block %d thread %d barrier@0x%x parity %u\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_a(bar, parity))
        if (++spins > (1u << 24)) lat_timeout(bar, parity);
}
__device__ __forceinline__ void tma_load_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------------ conditioning layout
// fp16, [N][L][tiles][8 warps][32 lanes][16 B]: the uint4 of thread (w, lane = 4 g + t) holds, as half2 pairs of channels
// (c, c+1), c = 8 w + 2 t:  .x = (row g, tanh c) .y = (row g+8, tanh c) .z = (row g, sigmoid R+c) .w = (row g+8, sigmoid R+c)
// i.e. exactly the accumulator fragment the thread adds it to.  Rows past the batch are zero.
__global__ void lat_cond_kernel(unsigned char* __restrict__ dst, const float* __restrict__ src, int first_sample, int nsamples, int L, int B, int ntiles)
{
    const size_t total = (size_t)nsamples * L * ntiles * 256;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int th = (int)(i & 255), w = th >> 5, lane = th & 31, g = lane >> 2, t = lane & 3;
        const size_t slt = i >> 8;                       // (s * L + l) * ntiles + tile
        const int tile = (int)(slt % ntiles);
        const size_t sl = slt / ntiles;                  // s * L + l
        const int c = 8 * w + 2 * t;
        uint32_t o[4] = {0, 0, 0, 0};
#pragma unroll
        for (int hi = 0; hi < 2; hi++) {
            const int b = tile * TU + g + 8 * hi;
            if (b < B) {
                const float* row = src + (sl * B + b) * 128;
                const float2 ft = *reinterpret_cast<const float2*>(row + c), fs = *reinterpret_cast<const float2*>(row + 64 + c);
                __half2 a = __floats2half2_rn(ft.x, ft.y), s = __floats2half2_rn(fs.x, fs.y);
                o[hi] = *reinterpret_cast<uint32_t*>(&a);
                o[2 + hi] = *reinterpret_cast<uint32_t*>(&s);
            }
        }
        const size_t off = (((size_t)first_sample * L * ntiles) + slt) * 4096 + (size_t)th * 16;
        *reinterpret_cast<uint4*>(dst + off) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ------------------------------------------------------------------------------------------------ pack
// blob (fp16, column-major M x K matrices as uploaded) -> weight image in mma.m16n8k16 B-fragment order.
// Element (n-tile nt, k-step pair jp, lane = 4 g + t) of a matrix W[M][K] is the uint4
//   { W[8nt+g][32jp+2t .. +1], W[8nt+g][32jp+8+2t ..], W[8nt+g][32jp+16+2t ..], W[8nt+g][32jp+24+2t ..] }
// = (b0, b1) of k-step 2jp and (b0, b1) of k-step 2jp+1.
__device__ __forceinline__ void frag_pos(int row, int k, int njp, size_t& byte_off)
{
    const int nt = row >> 3, g = row & 7, jp = k >> 5, kk = k & 31, comp = kk >> 3, t = (kk & 7) >> 1, e = kk & 1;
    byte_off = ((size_t)(nt * njp + jp) * 32 + (g * 4 + t)) * 16 + comp * 4 + e * 2;
}
__global__ void lat_pack_kernel(WnParams p, unsigned char* __restrict__ img, LatImage im)
{
    const int S = p.S, L = p.L;
    const __half* Wprev = static_cast<const __half*>(p.Wprev);
    const __half* Wcur = static_cast<const __half*>(p.Wcur);
    const __half* Wres = static_cast<const __half*>(p.Wres);
    const __half* Wskip = static_cast<const __half*>(p.Wskip);
    const __half* Wzs = static_cast<const __half*>(p.Wzs);
    const __half* Wza = static_cast<const __half*>(p.Wza);
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    const size_t g0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto put = [&](size_t base, int row, int k, int njp, __half v) {
        size_t o;
        frag_pos(row, k, njp, o);
        *reinterpret_cast<__half*>(img + base + o) = v;
    };
    for (size_t i = g0; i < (size_t)L * 128 * 64; i += gstride) {
        const int l = (int)(i / (128 * 64)), c = (int)(i % (128 * 64)) / 64, k = (int)(i % 64);
        const size_t lb = (size_t)l * im.layer_bytes;
        put(lb, c, k, 2, Wprev[(size_t)l * 128 * 64 + c + (size_t)k * 128]);
        put(lb + 16384, c, k, 2, Wcur[(size_t)l * 128 * 64 + c + (size_t)k * 128]);
        if (c < 64) put(lb + 32768, c, k, 2, Wres[(size_t)l * 64 * 64 + c + (size_t)k * 64]);
    }
    for (size_t i = g0; i < (size_t)L * S * 64; i += gstride) {
        const int l = (int)(i / ((size_t)S * 64)), s = (int)((i / 64) % S), k = (int)(i % 64);
        put((size_t)l * im.layer_bytes + 40960, s, k, 2, Wskip[(size_t)l * S * 64 + s + (size_t)k * S]);
    }
    // output matrices: stage load q holds k-step pairs [q OJP, (q+1) OJP) of all 32 n-tiles
    const int ojp = S == 256 ? 4 : 2;
    const size_t oload = (size_t)32 * ojp * 512;
    for (size_t i = g0; i < (size_t)A * S; i += gstride) {
        const int a = (int)(i / S), s = (int)(i % S);
        const int jp = s >> 5, q = jp / ojp;
        put(im.off_zs + (size_t)q * oload, a, (s & 31) + 32 * (jp % ojp), ojp, Wzs[a + (size_t)s * A]);
    }
    for (size_t i = g0; i < (size_t)A * A; i += gstride) {
        const int a = (int)(i / A), z = (int)(i % A);
        const int jp = z >> 5, q = jp / ojp;
        put(im.off_za + (size_t)q * oload, a, (z & 31) + 32 * (jp % ojp), ojp, Wza[a + (size_t)z * A]);
    }
    float* bias = reinterpret_cast<float*>(img + im.off_bias);
    const __half* Bh = static_cast<const __half*>(p.Bh);
    const __half* Bres = static_cast<const __half*>(p.Bres);
    const __half* Bskip = static_cast<const __half*>(p.Bskip);
    for (size_t i = g0; i < (size_t)L * 256; i += gstride) {
        const int l = (int)(i / 256), w = (int)(i % 256) / 32, t = (int)(i % 32) / 8, e = (int)(i % 8);
        const int c = 8 * w + 2 * t + (e & 1);
        float v = 0.f;
        if (e < 2) v = __half2float(Bh[(size_t)l * 128 + c]);
        else if (e < 4) v = __half2float(Bh[(size_t)l * 128 + 64 + c]);
        else if (e < 6) v = __half2float(Bres[(size_t)l * 64 + c]);
        bias[im.b_layer + i] = v;
    }
    for (size_t s = g0; s < (size_t)S; s += gstride) {
        float acc = 0.f;
        for (int l = 0; l < L; l++) { acc += __half2float(Bskip[(size_t)l * S + s]); bias[im.b_skpre + (size_t)l * S + s] = acc; }
    }
    for (size_t i = g0; i < (size_t)A; i += gstride) {
        bias[im.b_bzs + i] = __half2float(static_cast<const __half*>(p.Bzs)[i]);
        bias[im.b_bza + i] = __half2float(static_cast<const __half*>(p.Bza)[i]);
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int S>
__global__ void __launch_bounds__(NT, 1) wn_lat_kernel(const WnParams p, const unsigned char* __restrict__ img, const int ntiles_alloc)
{
    using C = Cfg<S>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const uint32_t sm = smem_u32(smem_raw);
    const int L = p.L, B = p.B;
    const LatImage im = lat_image(S, L);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x;
    const int slots = p.maxDil + 1;
    const int t_begin = p.init_sample, t_end = p.init_sample + p.count;
    const float* gbias = reinterpret_cast<const float*>(img + im.off_bias);

    const uint32_t s_full = sm + C::O_BAR;            // [2 stages][4]
    const uint32_t s_empty = s_full + 8 * 8;          // [2]
    int* dil = reinterpret_cast<int*>(smem_raw + C::O_DIL);
    int* ys = reinterpret_cast<int*>(smem_raw + C::O_YS);      // [TU] current index, [TU] previous index
    float* s_bout = reinterpret_cast<float*>(smem_raw + C::O_BOUT);

    // debug timeline: role 0 = compute thread 0, role 2 = producer; words (tag << 48 | clock)
    unsigned long long* trc = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
    int trn = 0;
    const int tr_t = p.trace_t & 0xFFFF;
#define TRACE(role, tag) do { if (trc && t == tr_t && trn < 1023) trc[(role) * 1024 + trn++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); } while (0)

    if (tid == 0) {
        for (int i = 0; i < 8; i++) mbar_init_a(s_full + 8 * i, 1);
        mbar_init_a(s_empty, NCW); mbar_init_a(s_empty + 8, NCW);
        fence_mbar_init();
        int d = 1;                                     // dilation of layer l (nv_wavenet.cuh:99-111): 1,2,4..maxDil,1,2,...
        for (int l = 0; l < L; l++) { dil[l] = d; d <<= 1; if (d > p.maxDil) d = 1; }
    }
    {   // embedding table of the current sample's index -> shared memory (rows padded to 33 words: gathers of different
        // rows fall into different banks); output-layer biases
        const uint32_t* ec = static_cast<const uint32_t*>(p.embCur);
        for (int i = tid; i < A * 32; i += NT) sts32(sm + C::O_EMB + ((i >> 5) * EROW + (i & 31)) * 4, ec[i]);
        for (int i = tid; i < S; i += NT) s_bout[i] = gbias[im.b_skpre + (size_t)(L - 1) * S + i];
        for (int i = tid; i < A; i += NT) { s_bout[S + i] = gbias[im.b_bzs + i]; s_bout[S + A + i] = gbias[im.b_bza + i]; }
        if (tid < TU) {
            const int b = tile * TU + tid;
            ys[tid] = b < B ? p.yCur[b] : 128;
            ys[TU + tid] = b < B ? p.yPrev[b] : 128;
        }
    }
    __syncthreads();

    if (warp == NCW) {
        // =============================================================== TMA producer (one lane)
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int t = t_begin; t < t_end; t++) {
                for (int l = 0; l < L; l++, cnt++) {
                    const uint32_t st = cnt & 1, dst = sm + C::O_RING + st * C::LAYER_BYTES, fb = s_full + st * 32;
                    mbar_wait_a(s_empty + 8 * st, ((cnt >> 1) & 1) ^ 1);
                    const unsigned char* src = img + (size_t)l * im.layer_bytes;
                    mbar_expect_a(fb, 16384);          tma_load_a(dst + C::W_PREV, src + C::W_PREV, 16384, fb);
                    mbar_expect_a(fb + 8, 16384);      tma_load_a(dst + C::W_CUR, src + C::W_CUR, 16384, fb + 8);
                    mbar_expect_a(fb + 16, 8192);      tma_load_a(dst + C::W_RES, src + C::W_RES, 8192, fb + 16);
                    mbar_expect_a(fb + 24, S * 128);   tma_load_a(dst + C::W_SKIP, src + C::W_SKIP, S * 128, fb + 24);
                    TRACE(2, 100 + l);
                }
                for (int q = 0; q < C::NQ_ZS + C::NQ_ZA; q++, cnt++) {
                    const uint32_t st = cnt & 1, dst = sm + C::O_RING + st * C::LAYER_BYTES, fb = s_full + st * 32;
                    mbar_wait_a(s_empty + 8 * st, ((cnt >> 1) & 1) ^ 1);
                    const unsigned char* src = q < C::NQ_ZS ? img + im.off_zs + (size_t)q * C::OLOAD : img + im.off_za + (size_t)(q - C::NQ_ZS) * C::OLOAD;
#pragma unroll
                    for (int k = 0; k < 4; k++) {      // quarter k = n-tiles [8k, 8k+8) = compute warps 2k, 2k+1
                        mbar_expect_a(fb + 8 * k, C::OLOAD / 4);
                        tma_load_a(dst + k * (C::OLOAD / 4), src + k * (C::OLOAD / 4), C::OLOAD / 4, fb + 8 * k);
                    }
                    TRACE(2, 200 + q);
                }
            }
        }
    } else {
        // =============================================================== compute warps
        const int w = warp, g = lane >> 2, t4 = lane & 3;
        const int b0 = tile * TU + g, b1 = b0 + 8;
        const bool v0 = b0 < B, v1 = b1 < B;
        const unsigned char* gcond = static_cast<const unsigned char*>(p.Lh) + (size_t)tile * 4096 + (size_t)(w * 32 + lane) * 16;
        unsigned char* gring = static_cast<unsigned char*>(p.ring) + (size_t)tile * 2048 + (size_t)lane * 16;
        const float* gbl = gbias + im.b_layer + (size_t)(w * 4 + t4) * 8;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        auto cond_ld = [&](int t, int l) -> uint4 {
            if (t >= t_end) return zero4;
            return ldg_nc_v4(gcond + ((size_t)t * L + l) * ntiles_alloc * 4096);
        };
        auto ring_ptr = [&](int t, int l) -> unsigned char* { return gring + ((size_t)(t % slots) * L + l) * ntiles_alloc * 2048; };
        auto prev_ld = [&](uint32_t (&x)[4][4], int t, int l) {
            const int d = dil[l];
            if (t >= t_end || t < d) {
#pragma unroll
                for (int j = 0; j < 4; j++) { x[j][0] = x[j][1] = x[j][2] = x[j][3] = 0; }
                return;
            }
            const unsigned char* src = ring_ptr(t - d, l);
#pragma unroll
            for (int j = 0; j < 4; j++) { const uint4 v = ldg_cg_v4(src + j * 512); x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w; }
        };
        const int jw = w >> 1, hw = w & 1;             // this warp's 8 channels = k-step jw, half hw of an activation tile
        const uint32_t xchg = (uint32_t)(jw * 512 + lane * 16 + hw * 8);

        uint32_t xa[4][4], xp[4][4], xpn[4][4];
        float xres[4];
        float sk[C::NSK][4];
#pragma unroll
        for (int i = 0; i < C::NSK; i++) sk[i][0] = sk[i][1] = sk[i][2] = sk[i][3] = 0.f;

        // previous-index embedding rows of the first sample -> epbuf (A-fragment order)
        if (w < 4) {
            const uint32_t* ep = static_cast<const uint32_t*>(p.embPrev);
            const int yp0 = ys[TU + g], yp1 = ys[TU + g + 8];
            uint4 v;
            v.x = ep[yp0 * 32 + 8 * w + t4]; v.y = ep[yp1 * 32 + 8 * w + t4];
            v.z = ep[yp0 * 32 + 8 * w + 4 + t4]; v.w = ep[yp1 * 32 + 8 * w + 4 + t4];
            sts128(sm + C::O_EPBUF + w * 512 + lane * 16, v);
        }
        // prefetch pipeline: conditioning two layers ahead, biases and history one layer ahead
        uint4 cd0 = cond_ld(t_begin, 0), cd1 = L > 1 ? cond_ld(t_begin, 1) : cond_ld(t_begin + 1, 0), cd2;
        float4 bs0 = *reinterpret_cast<const float4*>(gbl), bs1;
        float2 br0 = *reinterpret_cast<const float2*>(gbl + 4), br1;
        prev_ld(xp, t_begin, 0);
        bar_compute();

        uint32_t cnt = 0;
        for (int t = t_begin; t < t_end; t++) {
            const bool last = p.dump && t == t_end - 1;
            // ---------------- embedding (reference.cpp:42-57): x0 = [tanh](embPrev[yPrev] + embCur[yCur]); every warp builds
            // the whole A-fragment set redundantly from shared memory -- no exchange, no barrier
            const int yc0 = ys[g], yc1 = ys[g + 8];
            const float sel0 = (2 * w + 0 + tile * TU) < B ? p.sel[(size_t)t * B + tile * TU + 2 * w] : 0.5f;
            const float sel1 = (2 * w + 1 + tile * TU) < B ? p.sel[(size_t)t * B + tile * TU + 2 * w + 1] : 0.5f;
            if (tid == 0) TRACE(0, 1);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint4 pv = lds128(sm + C::O_EPBUF + j * 512 + lane * 16);
                const uint32_t pvv[4] = {pv.x, pv.y, pv.z, pv.w};
                uint32_t cv[4];
                cv[0] = lds32(sm + C::O_EMB + (yc0 * EROW + 8 * j + t4) * 4);
                cv[1] = lds32(sm + C::O_EMB + (yc1 * EROW + 8 * j + t4) * 4);
                cv[2] = lds32(sm + C::O_EMB + (yc0 * EROW + 8 * j + 4 + t4) * 4);
                cv[3] = lds32(sm + C::O_EMB + (yc1 * EROW + 8 * j + 4 + t4) * 4);
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const float2 a = unpack_h2(pvv[r]), c = unpack_h2(cv[r]);
                    float e0 = a.x + c.x, e1 = a.y + c.y;
                    if (p.tanhEmbed) { e0 = wn::tanhf_fast(e0); e1 = wn::tanhf_fast(e1); }
                    xa[j][r] = pack_h2(e0, e1);
                    if (j == jw && r == 2 * hw) { xres[0] = e0; xres[1] = e1; }
                    if (j == jw && r == 2 * hw + 1) { xres[2] = e0; xres[3] = e1; }
                }
            }
            // history of layer 0, and the previous-index rows of the NEXT sample (= this sample's current index)
            uint4 epn = zero4;
            if (w < 4) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (j == w) stg_v4(ring_ptr(t, 0) + j * 512, make_uint4(xa[j][0], xa[j][1], xa[j][2], xa[j][3]));
                const uint32_t* ep = static_cast<const uint32_t*>(p.embPrev);
                epn.x = __ldg(ep + yc0 * 32 + 8 * w + t4); epn.y = __ldg(ep + yc1 * 32 + 8 * w + t4);
                epn.z = __ldg(ep + yc0 * 32 + 8 * w + 4 + t4); epn.w = __ldg(ep + yc1 * 32 + 8 * w + 4 + t4);
            }
            if (tid == 0) TRACE(0, 2);
            const int lep = L > 2 ? 2 : L - 1;          // layer at which the next sample's previous-index rows are parked

            for (int l = 0; l < L; l++, cnt++) {
                const uint32_t st = sm + C::O_RING + (cnt & 1) * C::LAYER_BYTES, fb = s_full + (cnt & 1) * 32, ph = (cnt >> 1) & 1;
                // ---- prefetches
                int t1 = t, l1 = l + 1; if (l1 == L) { l1 = 0; t1 = t + 1; }
                int t2 = t1, l2 = l1 + 1; if (l2 == L) { l2 = 0; t2 = t1 + 1; }
                prev_ld(xpn, t1, l1);
                cd2 = cond_ld(t2, l2);
                bs1 = *reinterpret_cast<const float4*>(gbl + (size_t)l1 * 256);
                br1 = *reinterpret_cast<const float2*>(gbl + (size_t)l1 * 256 + 4);
                // ---- a = Wcur.x + Wprev.x[t-d] + (Bh + Lh)   (nv_wavenet.cuh:131-157)
                float acc[2][4];
                {
                    const float2 c0 = unpack_h2(cd0.x), c1 = unpack_h2(cd0.y), c2 = unpack_h2(cd0.z), c3 = unpack_h2(cd0.w);
                    acc[0][0] = bs0.x + c0.x; acc[0][1] = bs0.y + c0.y; acc[0][2] = bs0.x + c1.x; acc[0][3] = bs0.y + c1.y;
                    acc[1][0] = bs0.z + c2.x; acc[1][1] = bs0.w + c2.y; acc[1][2] = bs0.z + c3.x; acc[1][3] = bs0.w + c3.y;
                }
                mbar_wait_a(fb + 8, ph);
#pragma unroll
                for (int jp = 0; jp < 2; jp++) {
                    const uint4 bt = lds128(st + C::W_CUR + (w * 2 + jp) * 512 + lane * 16);
                    const uint4 bg = lds128(st + C::W_CUR + ((8 + w) * 2 + jp) * 512 + lane * 16);
                    hmma(acc[0], xa[2 * jp], bt.x, bt.y); hmma(acc[1], xa[2 * jp], bg.x, bg.y);
                    hmma(acc[0], xa[2 * jp + 1], bt.z, bt.w); hmma(acc[1], xa[2 * jp + 1], bg.z, bg.w);
                }
                mbar_wait_a(fb, ph);
#pragma unroll
                for (int jp = 0; jp < 2; jp++) {
                    const uint4 bt = lds128(st + C::W_PREV + (w * 2 + jp) * 512 + lane * 16);
                    const uint4 bg = lds128(st + C::W_PREV + ((8 + w) * 2 + jp) * 512 + lane * 16);
                    hmma(acc[0], xp[2 * jp], bt.x, bt.y); hmma(acc[1], xp[2 * jp], bg.x, bg.y);
                    hmma(acc[0], xp[2 * jp + 1], bt.z, bt.w); hmma(acc[1], xp[2 * jp + 1], bg.z, bg.w);
                }
                // ---- h = tanh(a[:R]) * sigmoid(a[R:])   (packed fp16 MUFU; sigmoid(x) = 0.5 tanh(x/2) + 0.5)
                const __half2 half = __floats2half2_rn(0.5f, 0.5f);
                const __half2 tg0 = wn::tanh_h2(h2(pack_h2(acc[0][0], acc[0][1]))), tg1 = wn::tanh_h2(h2(pack_h2(acc[0][2], acc[0][3])));
                const __half2 sg0 = __hfma2(wn::tanh_h2(h2(pack_h2(0.5f * acc[1][0], 0.5f * acc[1][1]))), half, half);
                const __half2 sg1 = __hfma2(wn::tanh_h2(h2(pack_h2(0.5f * acc[1][2], 0.5f * acc[1][3]))), half, half);
                sts64(sm + C::O_HBUF + xchg, u32(__hmul2(tg0, sg0)), u32(__hmul2(tg1, sg1)));
                bar_compute();
                if (l == lep && w < 4) sts128(sm + C::O_EPBUF + w * 512 + lane * 16, epn);   // every warp has read the old rows long ago
                uint32_t ha[4][4];
#pragma unroll
                for (int j = 0; j < 4; j++) load_a(ha[j], sm + C::O_HBUF + j * 512 + lane * 16);
                if (tid == 0) TRACE(0, 10);
                // ---- x' = Wres.h + Bres + x   (nv_wavenet.cuh:185-207)
                float ra[4] = {br0.x + xres[0], br0.y + xres[1], br0.x + xres[2], br0.y + xres[3]};
                mbar_wait_a(fb + 16, ph);
#pragma unroll
                for (int jp = 0; jp < 2; jp++) {
                    const uint4 bw = lds128(st + C::W_RES + (w * 2 + jp) * 512 + lane * 16);
                    hmma(ra, ha[2 * jp], bw.x, bw.y);
                    hmma(ra, ha[2 * jp + 1], bw.z, bw.w);
                }
                xres[0] = ra[0]; xres[1] = ra[1]; xres[2] = ra[2]; xres[3] = ra[3];
                if (l + 1 < L) {
                    const uint32_t x01 = pack_h2(ra[0], ra[1]), x23 = pack_h2(ra[2], ra[3]);
                    sts64(sm + C::O_XBUF + xchg, x01, x23);
                    stg_v2(ring_ptr(t, l + 1) + jw * 512 + hw * 8, x01, x23);
                }
                if (last) {
                    const int c = 8 * w + 2 * t4;
                    if (v0) { p.xtOut[((size_t)l * B + b0) * R + c] = ra[0]; p.xtOut[((size_t)l * B + b0) * R + c + 1] = ra[1]; }
                    if (v1) { p.xtOut[((size_t)l * B + b1) * R + c] = ra[2]; p.xtOut[((size_t)l * B + b1) * R + c + 1] = ra[3]; }
                }
                // ---- skip += Wskip.h   (biases are added once, after the last layer)
                mbar_wait_a(fb + 24, ph);
#pragma unroll
                for (int i = 0; i < C::NSK; i++) {
#pragma unroll
                    for (int jp = 0; jp < 2; jp++) {
                        const uint4 bw = lds128(st + C::W_SKIP + ((w * C::NSK + i) * 2 + jp) * 512 + lane * 16);
                        hmma(sk[i], ha[2 * jp], bw.x, bw.y);
                        hmma(sk[i], ha[2 * jp + 1], bw.z, bw.w);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(s_empty + 8 * (cnt & 1));
                if (last) {
                    const float* pre = gbias + im.b_skpre + (size_t)l * S;
#pragma unroll
                    for (int i = 0; i < C::NSK; i++) {
                        const int c = 8 * (w * C::NSK + i) + 2 * t4;
                        float o0 = sk[i][0] + pre[c], o1 = sk[i][1] + pre[c + 1], o2 = sk[i][2] + pre[c], o3 = sk[i][3] + pre[c + 1];
                        if (l == L - 1) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); o2 = fmaxf(o2, 0.f); o3 = fmaxf(o3, 0.f); }
                        if (v0) { p.skipOut[((size_t)l * B + b0) * S + c] = o0; p.skipOut[((size_t)l * B + b0) * S + c + 1] = o1; }
                        if (v1) { p.skipOut[((size_t)l * B + b1) * S + c] = o2; p.skipOut[((size_t)l * B + b1) * S + c + 1] = o3; }
                    }
                }
                if (l + 1 < L) {
                    bar_compute();
#pragma unroll
                    for (int j = 0; j < 4; j++) load_a(xa[j], sm + C::O_XBUF + j * 512 + lane * 16);
                }
                if (tid == 0) TRACE(0, 11);
#pragma unroll
                for (int j = 0; j < 4; j++) { xp[j][0] = xpn[j][0]; xp[j][1] = xpn[j][1]; xp[j][2] = xpn[j][2]; xp[j][3] = xpn[j][3]; }
                cd0 = cd1; cd1 = cd2; bs0 = bs1; br0 = br1;
            }

            // ---------------- relu(skip + bias) -> Zs -> Za   (reference.cpp:93-104)
#pragma unroll
            for (int i = 0; i < C::NSK; i++) {
                const int nt = w * C::NSK + i, c = 8 * nt + 2 * t4;
                const float b0f = s_bout[c], b1f = s_bout[c + 1];
                sts64(sm + C::O_OB0 + (nt >> 1) * 512 + lane * 16 + (nt & 1) * 8,
                      pack_h2(fmaxf(sk[i][0] + b0f, 0.f), fmaxf(sk[i][1] + b1f, 0.f)), pack_h2(fmaxf(sk[i][2] + b0f, 0.f), fmaxf(sk[i][3] + b1f, 0.f)));
                sk[i][0] = sk[i][1] = sk[i][2] = sk[i][3] = 0.f;
            }
            bar_compute();
            if (tid == 0) TRACE(0, 20);
            float zz[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = 32 * w + 8 * i + 2 * t4;
                zz[i][0] = zz[i][2] = s_bout[S + c]; zz[i][1] = zz[i][3] = s_bout[S + c + 1];
            }
            for (int q = 0; q < C::NQ_ZS; q++, cnt++) {
                const uint32_t st = sm + C::O_RING + (cnt & 1) * C::LAYER_BYTES, fb = s_full + (cnt & 1) * 32, ph = (cnt >> 1) & 1;
                mbar_wait_a(fb + 8 * (w >> 1), ph);
#pragma unroll
                for (int jp = 0; jp < C::OJP; jp++) {
                    uint32_t a0[4], a1[4];
                    load_a(a0, sm + C::O_OB0 + ((q * C::OJP + jp) * 2) * 512 + lane * 16);
                    load_a(a1, sm + C::O_OB0 + ((q * C::OJP + jp) * 2 + 1) * 512 + lane * 16);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 bw = lds128(st + ((4 * w + i) * C::OJP + jp) * 512 + lane * 16);
                        hmma(zz[i], a0, bw.x, bw.y);
                        hmma(zz[i], a1, bw.z, bw.w);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(s_empty + 8 * (cnt & 1));
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int nt = 4 * w + i, c = 8 * nt + 2 * t4;
                const float z0 = fmaxf(zz[i][0], 0.f), z1 = fmaxf(zz[i][1], 0.f), z2 = fmaxf(zz[i][2], 0.f), z3 = fmaxf(zz[i][3], 0.f);
                sts64(sm + C::O_OB1 + (nt >> 1) * 512 + lane * 16 + (nt & 1) * 8, pack_h2(z0, z1), pack_h2(z2, z3));
                if (last) {
                    if (v0) { p.Zs[(size_t)b0 * A + c] = z0; p.Zs[(size_t)b0 * A + c + 1] = z1; }
                    if (v1) { p.Zs[(size_t)b1 * A + c] = z2; p.Zs[(size_t)b1 * A + c + 1] = z3; }
                }
                zz[i][0] = zz[i][2] = s_bout[S + A + c]; zz[i][1] = zz[i][3] = s_bout[S + A + c + 1];
            }
            bar_compute();
            if (tid == 0) TRACE(0, 21);
            for (int q = 0; q < C::NQ_ZA; q++, cnt++) {
                const uint32_t st = sm + C::O_RING + (cnt & 1) * C::LAYER_BYTES, fb = s_full + (cnt & 1) * 32, ph = (cnt >> 1) & 1;
                mbar_wait_a(fb + 8 * (w >> 1), ph);
#pragma unroll
                for (int jp = 0; jp < C::OJP; jp++) {
                    uint32_t a0[4], a1[4];
                    load_a(a0, sm + C::O_OB1 + ((q * C::OJP + jp) * 2) * 512 + lane * 16);
                    load_a(a1, sm + C::O_OB1 + ((q * C::OJP + jp) * 2 + 1) * 512 + lane * 16);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint4 bw = lds128(st + ((4 * w + i) * C::OJP + jp) * 512 + lane * 16);
                        hmma(zz[i], a0, bw.x, bw.y);
                        hmma(zz[i], a1, bw.z, bw.w);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive_a(s_empty + 8 * (cnt & 1));
            }
            // logits (fp32) -> transposed buffer: row = utterance, 256 contiguous channels
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int c = 32 * w + 8 * i + 2 * t4;
                sts64(sm + C::O_LBUF + (g * LROW + c) * 4, __float_as_uint(zz[i][0]), __float_as_uint(zz[i][1]));
                sts64(sm + C::O_LBUF + ((g + 8) * LROW + c) * 4, __float_as_uint(zz[i][2]), __float_as_uint(zz[i][3]));
                if (last) {
                    if (v0) { p.Za[(size_t)b0 * A + c] = zz[i][0]; p.Za[(size_t)b0 * A + c + 1] = zz[i][1]; }
                    if (v1) { p.Za[(size_t)b1 * A + c] = zz[i][2]; p.Za[(size_t)b1 * A + c + 1] = zz[i][3]; }
                }
            }
            bar_compute();
            if (tid == 0) TRACE(0, 22);
            // ---------------- softmax + categorical sample (matrix.cpp:167-183, reference.cpp:106-121): warp w serves
            // utterances 2w and 2w+1; lane holds 8 consecutive classes of each
            {
                float e[2][8], m[2] = {0.f, 0.f};                          // the reference starts the max at 0 (matrix.cpp:171)
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const uint4 u0 = lds128(sm + C::O_LBUF + ((2 * w + r) * LROW + 8 * lane) * 4), u1 = lds128(sm + C::O_LBUF + ((2 * w + r) * LROW + 8 * lane + 4) * 4);
                    e[r][0] = __uint_as_float(u0.x); e[r][1] = __uint_as_float(u0.y); e[r][2] = __uint_as_float(u0.z); e[r][3] = __uint_as_float(u0.w);
                    e[r][4] = __uint_as_float(u1.x); e[r][5] = __uint_as_float(u1.y); e[r][6] = __uint_as_float(u1.z); e[r][7] = __uint_as_float(u1.w);
#pragma unroll
                    for (int k = 0; k < 8; k++) m[r] = fmaxf(m[r], e[r][k]);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    m[0] = fmaxf(m[0], __shfl_xor_sync(0xffffffffu, m[0], o));
                    m[1] = fmaxf(m[1], __shfl_xor_sync(0xffffffffu, m[1], o));
                }
                float incl[2];
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const float ms = m[r] * 1.4426950408889634f;
                    float run = 0.f;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float ev = wn::exp2f_fast(fmaf(e[r][k], 1.4426950408889634f, -ms));
                        run += ev;
                        e[r][k] = run;                                     // inclusive running sum inside the lane
                    }
                    incl[r] = run;
                }
                const float tot_lane[2] = {incl[0], incl[1]};
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float a0 = __shfl_up_sync(0xffffffffu, incl[0], o), a1 = __shfl_up_sync(0xffffffffu, incl[1], o);
                    if (lane >= o) { incl[0] += a0; incl[1] += a1; }
                }
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const float total = __shfl_sync(0xffffffffu, incl[r], 31);
                    const float excl = incl[r] - tot_lane[r];
                    const float target = (r == 0 ? sel0 : sel1) * total;
                    int cntk = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++) cntk += (target < excl + e[r][k]) ? 0 : 1;
                    const unsigned ball = __ballot_sync(0xffffffffu, target < incl[r]);
                    const int lf = ball ? __ffs(ball) - 1 : 31;
                    const int ck = __shfl_sync(0xffffffffu, cntk, lf);
                    const int y = ball ? 8 * lf + (ck < 7 ? ck : 7) : A - 1;
                    const int b = tile * TU + 2 * w + r;
                    if (last && b < B) {
                        const float inv = 1.f / total;
                        float prevv = 0.f;
#pragma unroll
                        for (int k = 0; k < 8; k++) { p.P[(size_t)b * A + 8 * lane + k] = (e[r][k] - prevv) * inv; prevv = e[r][k]; }
                    }
                    if (lane == 0) {
                        int fbk = y;
                        if (b < B) {
                            p.yOut[(size_t)b * p.N + t] = y;
                            if (p.forced) fbk = p.forced[(size_t)b * p.N + t];
                        } else fbk = 128;
                        ys[TU + 2 * w + r] = ys[2 * w + r];
                        ys[2 * w + r] = fbk;
                    }
                }
            }
            bar_compute();
            if (tid == 0) TRACE(0, 23);
        }
        if (tid < TU && tile * TU + tid < B) { p.yCur[tile * TU + tid] = ys[tid]; p.yPrev[tile * TU + tid] = ys[TU + tid]; }
    }
#undef TRACE
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host side
bool wn_lat_supported(int R_, int S, int A_, int L)
{
    return R_ == R && A_ == A && (S == 128 || S == 256) && L >= 1 && L <= MAXL;
}
int wn_lat_tiles(int B) { return (B + TU - 1) / TU; }
size_t wn_lat_image_bytes(int S, int L) { return lat_image(S, L).total; }
size_t wn_lat_ring_bytes(int L, int maxDil, int B) { return (size_t)(maxDil + 1) * L * wn_lat_tiles(B) * 2048; }
size_t wn_lat_cond_bytes(int L, int B, int N) { return (size_t)N * L * wn_lat_tiles(B) * 4096; }

cudaError_t wn_lat_cond_convert(void* dst, const float* src_dev, int first_sample, int nsamples, int L, int B, cudaStream_t stream)
{
    if (nsamples <= 0) return cudaSuccess;
    const int ntiles = wn_lat_tiles(B);
    const size_t total = (size_t)nsamples * L * ntiles * 256;
    size_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    lat_cond_kernel<<<(unsigned)blocks, 256, 0, stream>>>(static_cast<unsigned char*>(dst), src_dev, first_sample, nsamples, L, B, ntiles);
    return cudaGetLastError();
}

cudaError_t wn_lat_pack(void* image, const WnParams& p, cudaStream_t stream)
{
    const LatImage im = lat_image(p.S, p.L);
    cudaError_t e = cudaMemsetAsync(image, 0, im.total, stream);
    if (e != cudaSuccess) return e;
    lat_pack_kernel<<<296, 256, 0, stream>>>(p, static_cast<unsigned char*>(image), im);
    return cudaGetLastError();
}

// p.B = utterances of this run; engine_B = batch size the conditioning store / history ring were laid out for
cudaError_t wn_launch_lat(const WnParams& p, const void* image, int engine_B, cudaStream_t stream, WnLaunchInfo* info)
{
    const int grid = wn_lat_tiles(p.B), ntiles_alloc = wn_lat_tiles(engine_B);
    const unsigned char* im8 = static_cast<const unsigned char*>(image);
    cudaError_t e;
    size_t smem;
    if (p.S == 256) {
        smem = Cfg<256>::SMEM;
        e = cudaFuncSetAttribute(wn_lat_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wn_lat_kernel<256><<<grid, NT, smem, stream>>>(p, im8, ntiles_alloc);
    } else if (p.S == 128) {
        smem = Cfg<128>::SMEM;
        e = cudaFuncSetAttribute(wn_lat_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wn_lat_kernel<128><<<grid, NT, smem, stream>>>(p, im8, ntiles_alloc);
    } else {
        return cudaErrorInvalidValue;
    }
    if (info) { info->kernel = 18; info->grid = grid; info->block = NT; info->smem_bytes = (int)smem; info->batch_per_cta = TU; info->cluster = 1; }
    return cudaGetLastError();
}
