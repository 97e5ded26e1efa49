// wn_tc_kernel.cu -- fp16 tensor-core kernel of the WaveNet inference loop (sm_100a: tcgen05 + TMEM + bulk TMA).
//
// ONE persistent CTA per tile of up to 128 utterances runs the whole autoregressive loop for `count` samples:
//
//   warps 0-3  "epilogue": thread i owns utterance i of the tile (= TMEM lane i = row i of every activation tile).
//              embed -> per layer: [D1 + Bh + Lh -> tanh * sigmoid -> h tile] [Dx + Bres + x -> x tile, history ring]
//              -> relu(skip) tile -> relu(Zs) tile -> softmax + categorical sample, all thread-local
//              (the reference spreads these over CTAs/threads: nv_wavenet_persistent.cuh:223-462, softmax.cuh:36-191).
//   warp 4     lane 0: TMA producer.  Streams the pre-tiled fp16 weight image (and the x[t-d] history tiles) from
//              L2 through an NSTAGE x 16 KB shared-memory ring with cp.async.bulk + mbarrier complete_tx.
//   warp 5     lane 0: MMA issuer.  D[128 utterances x N channels] (fp32, TMEM) += X[128 x 64] . W[N x 64]^T with
//              tcgen05.mma (activations = A operand, weights = B operand, both K-major SWIZZLE_128B);
//              the skip accumulator lives in TMEM across all layers; tcgen05.commit signals the epilogue and
//              frees ring stages.
//
// Replaces nv_wavenet_persistent.cuh + matrix_math.cuh + softmax.cuh of the reference for T_data = half.
// Numerical contract (oracle/wavenet_oracle.c, WNO_PREC_FP16): weights, biases, embeddings, Lh and every GEMM
// input rounded to fp16; fp32 accumulation; residual stream, skip sum, softmax in fp32.
#include "wn_common.h"
#include "wn_math.cuh"
#include "wn_sm100.cuh"

namespace {

using namespace sm100;

constexpr int R = 64, A = 256;
constexpr int TILE = 16384;                 // [128 rows x 64 fp16] K-major SW128
constexpr int NT = 192;

struct TcImage {                            // byte offsets inside the packed image
    size_t layer_bytes, off_out, off_bias, total;
    size_t b_bh, b_bres, b_bskp, b_bzs, b_bza;     // float offsets inside the bias block
};
__host__ __device__ inline TcImage tc_image(int S, int L)
{
    TcImage im;
    im.layer_bytes = (size_t)TILE * 2 + TILE / 2 + (size_t)(S / 128) * TILE;
    im.off_out = (size_t)L * im.layer_bytes;
    im.off_bias = im.off_out + (size_t)(S / 64) * 2 * TILE + (size_t)(A / 64) * 2 * TILE;
    im.b_bh = 0;
    im.b_bres = im.b_bh + (size_t)L * 128;
    im.b_bskp = im.b_bres + (size_t)L * 64;
    im.b_bzs = im.b_bskp + (size_t)L * S;
    im.b_bza = im.b_bzs + A;
    im.total = im.off_bias + (im.b_bza + A) * sizeof(float);
    return im;
}

__host__ __device__ inline size_t tc_smem_bytes(int S, int L, int nstage)
{
    // 4 activation tiles + ring + biases (Bh, Bres, Bskip-sum, Bzs, Bza) + barriers
    return 1024 + 4 * (size_t)TILE + (size_t)nstage * TILE + ((size_t)L * 192 + S + 2 * A) * sizeof(float) + (2 * nstage + 8) * 8 + 16;
}

// ------------------------------------------------------------------------------------------------ pack
// blob (fp16, column-major matrices as uploaded) -> tiled / swizzled weight image + fp32 bias block
__global__ void tc_pack_kernel(WnParams p, unsigned char* __restrict__ img, TcImage im)
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
    auto put = [&](size_t chunk_off, int row, int k, __half v) { *reinterpret_cast<__half*>(img + chunk_off + sw128_offset(row, k)) = v; };
    // per layer
    for (size_t i = g0; i < (size_t)L * 128 * 64; i += gstride) {
        const int l = (int)(i / (128 * 64)), c = (int)(i % (128 * 64)) / 64, k = (int)(i % 64);
        const size_t lb = (size_t)l * im.layer_bytes;
        put(lb, c, k, Wprev[(size_t)l * 128 * 64 + c + (size_t)k * 128]);
        put(lb + TILE, c, k, Wcur[(size_t)l * 128 * 64 + c + (size_t)k * 128]);
        if (c < 64) put(lb + 2 * TILE, c, k, Wres[(size_t)l * 64 * 64 + c + (size_t)k * 64]);
    }
    for (size_t i = g0; i < (size_t)L * S * 64; i += gstride) {
        const int l = (int)(i / ((size_t)S * 64)), s = (int)((i / 64) % S), k = (int)(i % 64);
        put((size_t)l * im.layer_bytes + 2 * TILE + TILE / 2 + (size_t)(s / 128) * TILE, s % 128, k, Wskip[(size_t)l * S * 64 + s + (size_t)k * S]);
    }
    // output layers: chunk (kt, nh) = rows a in [128 nh, +128), k in [64 kt, +64)
    for (size_t i = g0; i < (size_t)A * S; i += gstride) {
        const int a = (int)(i / S), s = (int)(i % S);
        put(im.off_out + (size_t)((s / 64) * 2 + a / 128) * TILE, a % 128, s % 64, Wzs[a + (size_t)s * A]);
    }
    const size_t off_wza = im.off_out + (size_t)(S / 64) * 2 * TILE;
    for (size_t i = g0; i < (size_t)A * A; i += gstride) {
        const int a = (int)(i / A), z = (int)(i % A);
        put(off_wza + (size_t)((z / 64) * 2 + a / 128) * TILE, a % 128, z % 64, Wza[a + (size_t)z * A]);
    }
    // biases -> fp32; running prefix of the skip biases (the skip sum is kept in TMEM without biases)
    float* bias = reinterpret_cast<float*>(img + im.off_bias);
    const __half* Bh = static_cast<const __half*>(p.Bh);
    const __half* Bres = static_cast<const __half*>(p.Bres);
    const __half* Bskip = static_cast<const __half*>(p.Bskip);
    for (size_t i = g0; i < (size_t)L * 128; i += gstride) bias[im.b_bh + i] = __half2float(Bh[i]);
    for (size_t i = g0; i < (size_t)L * 64; i += gstride) bias[im.b_bres + i] = __half2float(Bres[i]);
    for (size_t s = g0; s < (size_t)S; s += gstride) {
        float acc = 0.f;
        for (int l = 0; l < L; l++) { acc += __half2float(Bskip[(size_t)l * S + s]); bias[im.b_bskp + (size_t)l * S + s] = acc; }
    }
    for (size_t i = g0; i < (size_t)A; i += gstride) {
        bias[im.b_bzs + i] = __half2float(static_cast<const __half*>(p.Bzs)[i]);
        bias[im.b_bza + i] = __half2float(static_cast<const __half*>(p.Bza)[i]);
    }
}

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t v)
{
    return __half22float2(*reinterpret_cast<__half2*>(&v));
}
// chunk q (16 bytes = 8 fp16) of row `row` inside a SW128 tile
__device__ __forceinline__ uint32_t chunk_off(int row, int q) { return (uint32_t)row * 128u + (uint32_t)((q ^ (row & 7)) << 4); }

template <int S>
__global__ void __launch_bounds__(NT, 1) wn_tc_kernel(const WnParams p, const unsigned char* __restrict__ img, const int nstage)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int L = p.L, B = p.B;
    const TcImage im = tc_image(S, L);

    unsigned char* t_xc = smem;                    // x tile of the current layer        (= BIG k-tile 0)
    unsigned char* t_h = smem + TILE;              // gated activation tile              (= BIG k-tile 1)
    unsigned char* t_big = smem;                   // [128 x 256] as 4 k-tiles: relu(skip) then relu(Zs)
    unsigned char* ring = smem + 4 * TILE;
    float* s_bh = reinterpret_cast<float*>(ring + (size_t)nstage * TILE);
    float* s_bres = s_bh + (size_t)L * 128;
    float* s_bsk = s_bres + (size_t)L * 64;
    float* s_bzs = s_bsk + S;
    float* s_bza = s_bzs + A;
    uint64_t* w_full = reinterpret_cast<uint64_t*>(s_bza + A);
    uint64_t* w_empty = w_full + nstage;
    uint64_t* epi_done = w_empty + nstage;
    uint64_t* d1_full = epi_done + 1;
    uint64_t* dx_full = epi_done + 2;
    uint64_t* skip_full = epi_done + 3;
    uint64_t* out_full = epi_done + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(epi_done + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, ntiles = gridDim.x;
    const int slots = p.maxDil + 1;
    const int t_begin = p.init_sample, t_end = p.init_sample + p.count;
    unsigned char* gring = static_cast<unsigned char*>(p.ring);
    auto ring_tile = [&](int t, int l) -> unsigned char* {
        return gring + (((size_t)(t % slots) * L + l) * ntiles + tile) * (size_t)TILE;
    };

    if (tid == 0) {
        for (int s = 0; s < nstage; s++) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        mbar_init(epi_done, 128);
        mbar_init(d1_full, 1); mbar_init(dx_full, 1); mbar_init(skip_full, 1); mbar_init(out_full, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(tmem_slot);
    {   // biases -> shared memory
        const float* gb = reinterpret_cast<const float*>(img + im.off_bias);
        for (int i = tid; i < L * 128; i += NT) s_bh[i] = gb[im.b_bh + i];
        for (int i = tid; i < L * 64; i += NT) s_bres[i] = gb[im.b_bres + i];
        for (int i = tid; i < S; i += NT) s_bsk[i] = gb[im.b_bskp + (size_t)(L - 1) * S + i];
        for (int i = tid; i < A; i += NT) { s_bzs[i] = gb[im.b_bzs + i]; s_bza[i] = gb[im.b_bza + i]; }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t D1 = tmem_base, DX = tmem_base + 128, DSKIP = tmem_base + 256, DZS = tmem_base + 256, DZA = tmem_base;

    // dilation of layer l (nv_wavenet.cuh:99-111): 1,2,4..maxDil,1,2,...
    int cyc = 0;
    for (int d = 1; d <= p.maxDil; d <<= 1) cyc++;
    auto dil = [&](int l) -> int { return 1 << (l % cyc); };

    // debug timeline: role r (0 epilogue thread 0, 1 MMA issuer, 2 producer) appends (tag << 48 | clock) words
    unsigned long long* trc = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
    int trn = 0;
#define TRACE(role, tag) do { if (trc && t == p.trace_t && trn < 1023) trc[(role) * 1024 + trn++] = ((unsigned long long)(tag) << 48) | (clock64() & 0xFFFFFFFFFFFFull); } while (0)

    if (warp == 4) {
        // =============================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t ph = 1;
            auto put = [&](const void* src, uint32_t bytes) {
                mbar_wait(&w_empty[stage], ph);
                mbar_arrive_expect_tx(&w_full[stage], bytes);
                tma_load_1d(ring + (size_t)stage * TILE, src, bytes, &w_full[stage]);
                if (++stage == nstage) { stage = 0; ph ^= 1; }
            };
            for (int t = t_begin; t < t_end; t++) {
                for (int l = 0; l < L; l++) {
                    const unsigned char* lw = img + (size_t)l * im.layer_bytes;
                    const int d = dil(l);
                    TRACE(2, 100 + l);
                    if (t >= d) { put(ring_tile(t - d, l), TILE); put(lw, TILE); }
                    put(lw + TILE, TILE);
                    put(lw + 2 * TILE, TILE / 2);
                    for (int c = 0; c < S / 128; c++) put(lw + 2 * TILE + TILE / 2 + (size_t)c * TILE, TILE);
                }
                const unsigned char* ow = img + im.off_out;
                for (int c = 0; c < (S / 64) * 2 + (A / 64) * 2; c++) put(ow + (size_t)c * TILE, TILE);
            }
        }
    } else if (warp == 5) {
        // =============================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc128 = make_idesc_f16(128, 128), idesc64 = make_idesc_f16(128, 64);
            int stage = 0;
            uint32_t ph_full = 0, ph_epi = 0;
            const uint32_t ring_a = smem_u32(ring), xc_a = smem_u32(t_xc), h_a = smem_u32(t_h), big_a = smem_u32(t_big);
            auto wait_stage = [&]() -> int {       // returns the stage index whose data has landed, advances
                mbar_wait(&w_full[stage], ph_full);
                const int s = stage;
                if (++stage == nstage) { stage = 0; ph_full ^= 1; }
                return s;
            };
            auto mma4 = [&](uint32_t a_addr, uint32_t b_addr, uint32_t d, uint32_t idesc, bool acc0) {
                const uint64_t da = make_desc_kmajor_sw128(a_addr), db = make_desc_kmajor_sw128(b_addr);
#pragma unroll
                for (int k = 0; k < 4; k++) umma_f16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (acc0 || k) ? 1u : 0u);
            };
            auto wait_epi = [&]() { mbar_wait(epi_done, ph_epi); ph_epi ^= 1; tc_fence_after_sync(); };
            auto prev = [&](int l) {               // D1 = Wprev . x[t-d]   (overwrites D1)
                const int sa = wait_stage(), sb = wait_stage();
                tc_fence_after_sync();
                mma4(ring_a + sa * TILE, ring_a + sb * TILE, D1, idesc128, false);
                umma_commit(&w_empty[sa]);
                umma_commit(&w_empty[sb]);
            };
            for (int t = t_begin; t < t_end; t++) {
                const bool dump = p.dump && (t == t_end - 1);
                for (int l = 0; l < L; l++) {
                    wait_epi();                                         // x_l tile ready (and, for l = 0, Dza consumed)
                    TRACE(1, 20);
                    const bool has_prev = t >= dil(l);
                    if (l == 0 && has_prev) prev(0);
                    {   // D1 += Wcur . x[t]
                        const int s = wait_stage();
                        tc_fence_after_sync();
                        mma4(xc_a, ring_a + s * TILE, D1, idesc128, has_prev);
                        umma_commit(&w_empty[s]);
                        umma_commit(d1_full);
                        TRACE(1, 21);
                    }
                    wait_epi();                                         // h tile ready, D1 consumed
                    TRACE(1, 22);
                    {   // Dx = Wres . h
                        const int s = wait_stage();
                        tc_fence_after_sync();
                        mma4(h_a, ring_a + s * TILE, DX, idesc64, false);
                        umma_commit(&w_empty[s]);
                        umma_commit(dx_full);
                        TRACE(1, 23);
                    }
                    for (int c = 0; c < S / 128; c++) {                 // Dskip (+)= Wskip . h   (accumulates over layers)
                        const int s = wait_stage();
                        tc_fence_after_sync();
                        mma4(h_a, ring_a + s * TILE, DSKIP + c * 128, idesc128, l > 0);
                        umma_commit(&w_empty[s]);
                    }
                    if (dump || l == L - 1) umma_commit(skip_full);
                    if (l + 1 < L && t >= dil(l + 1)) prev(l + 1);      // off the critical path
                    TRACE(1, 24);
                }
                wait_epi();                                             // relu(skip) tile ready
                for (int kt = 0; kt < S / 64; kt++)
                    for (int nh = 0; nh < 2; nh++) {
                        const int s = wait_stage();
                        tc_fence_after_sync();
                        mma4(big_a + kt * TILE, ring_a + s * TILE, DZS + nh * 128, idesc128, kt > 0);
                        umma_commit(&w_empty[s]);
                    }
                umma_commit(out_full);
                wait_epi();                                             // relu(Zs) tile ready
                for (int kt = 0; kt < A / 64; kt++)
                    for (int nh = 0; nh < 2; nh++) {
                        const int s = wait_stage();
                        tc_fence_after_sync();
                        mma4(big_a + kt * TILE, ring_a + s * TILE, DZA + nh * 128, idesc128, kt > 0);
                        umma_commit(&w_empty[s]);
                    }
                umma_commit(out_full);
            }
        }
    } else {
        // =============================================================== epilogue warps: one thread per utterance
        const int row = tid;
        const int b = tile * 128 + row;
        const bool valid = b < B;
        const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
        uint32_t ph_d1 = 0, ph_dx = 0, ph_skip = 0, ph_out = 0;
        const __half* embPrev = static_cast<const __half*>(p.embPrev);
        const __half* embCur = static_cast<const __half*>(p.embCur);
        const __half* Lh = static_cast<const __half*>(p.Lh);
        const float* gbias = reinterpret_cast<const float*>(img + im.off_bias);
        int yp = valid ? p.yPrev[b] : 0, yc = valid ? p.yCur[b] : 0;
        float x[R];                                   // residual stream of this utterance (fp32)
        uint32_t lh[64];                              // Lh[t][l][b][0:128] as packed fp16 pairs
        auto lh_ptr = [&](int t, int l) -> const uint4* {
            return reinterpret_cast<const uint4*>(Lh + (((size_t)t * L + l) * B + (valid ? b : 0)) * 128);
        };
        auto load_lh = [&](int t, int l) {
            const uint4* src = lh_ptr(t, l);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const uint4 v = ldg_nc_v4(src + i);
                lh[4 * i] = v.x; lh[4 * i + 1] = v.y; lh[4 * i + 2] = v.z; lh[4 * i + 3] = v.w;
            }
        };
        // History ring (global, read back d samples later by TMA): written AFTER the barrier arrival that publishes the
        // shared-memory tile, so that the proxy fence in front of that arrival never waits for global stores (or for the
        // conditioning loads below).  The stores are fenced by the next stage's fence.proxy.async, long before any TMA
        // read of them (>= one full sample later).
        auto store_history = [&](unsigned char* grow) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; j++) o[j] = pack_h2(x[8 * q + 2 * j], x[8 * q + 2 * j + 1]);
                *reinterpret_cast<uint4*>(grow + chunk_off(row, q)) = make_uint4(o[0], o[1], o[2], o[3]);
            }
        };
        if (t_begin < t_end) load_lh(t_begin, 0);

        for (int t = t_begin; t < t_end; t++) {
            const bool dump = p.dump && (t == t_end - 1);
            const float sel = valid ? __ldg(p.sel + (size_t)t * B + b) : 0.5f;
            // ---------------- embedding: x0 = tanh(embPrev[yPrev] + embCur[yCur])   (reference.cpp:42-57)
            {
                const uint4* ep = reinterpret_cast<const uint4*>(embPrev + (size_t)yp * R);
                const uint4* ec = reinterpret_cast<const uint4*>(embCur + (size_t)yc * R);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const uint4 a = __ldg(ep + q), c = __ldg(ec + q);
                    const uint32_t av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
                    uint32_t o[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float2 fa = unpack_h2(av[j]), fc = unpack_h2(cv[j]);
                        float e0 = fa.x + fc.x, e1 = fa.y + fc.y;
                        if (p.tanhEmbed) { e0 = wn::tanhf_fast(e0); e1 = wn::tanhf_fast(e1); }
                        if (!valid) { e0 = 0.f; e1 = 0.f; }
                        x[8 * q + 2 * j] = e0; x[8 * q + 2 * j + 1] = e1;
                        o[j] = pack_h2(e0, e1);
                    }
                    *reinterpret_cast<uint4*>(t_xc + chunk_off(row, q)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
                tc_fence_before_sync();
                fence_proxy_async();
                mbar_arrive(epi_done);                                  // x_0 ready
                if (tid == 0) TRACE(0, 1);
                store_history(ring_tile(t, 0));                         // off the critical path (see store_history)
            }

            for (int l = 0; l < L; l++) {
                // ---------------- gate: h = tanh(a[0:R]) * sigmoid(a[R:2R]),  a = D1 + Bh + Lh   (reference.cpp:67-80)
                mbar_wait(d1_full, ph_d1); ph_d1 ^= 1;
                tc_fence_after_sync();
                if (tid == 0) TRACE(0, 2);
                const float* bh = s_bh + (size_t)l * 128;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t ta[16], sa[16];
                    tmem_ld16(D1 + lane_off + 16 * q, ta);
                    tmem_ld16(D1 + lane_off + 64 + 16 * q, sa);
                    tmem_ld_wait();
                    uint32_t hp[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int r0 = 16 * q + j;
                        const float2 lt = unpack_h2(lh[r0 >> 1]), ls = unpack_h2(lh[32 + (r0 >> 1)]);
                        const float2 bt = *reinterpret_cast<const float2*>(bh + r0), bs = *reinterpret_cast<const float2*>(bh + 64 + r0);
                        const float a0 = __uint_as_float(ta[j]) + bt.x + lt.x, a1 = __uint_as_float(ta[j + 1]) + bt.y + lt.y;
                        const float g0 = __uint_as_float(sa[j]) + bs.x + ls.x, g1 = __uint_as_float(sa[j + 1]) + bs.y + ls.y;
                        const float h0 = wn::tanhf_fast(a0) * wn::sigmoidf_fast(g0);
                        const float h1 = wn::tanhf_fast(a1) * wn::sigmoidf_fast(g1);
                        hp[j >> 1] = pack_h2(h0, h1);
                    }
                    *reinterpret_cast<uint4*>(t_h + chunk_off(row, 2 * q)) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
                    *reinterpret_cast<uint4*>(t_h + chunk_off(row, 2 * q + 1)) = make_uint4(hp[4], hp[5], hp[6], hp[7]);
                }
                if (tid == 0) TRACE(0, 12);
                tc_fence_before_sync();
                fence_proxy_async();
                mbar_arrive(epi_done);                                  // h ready, D1 free
                if (tid == 0) TRACE(0, 3);
                // ---------------- residual: x += Dx + Bres   (reference.cpp:82-84)
                mbar_wait(dx_full, ph_dx); ph_dx ^= 1;
                tc_fence_after_sync();
                if (tid == 0) TRACE(0, 4);
                const float* br = s_bres + (size_t)l * 64;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    uint32_t v[16];
                    tmem_ld16(DX + lane_off + 16 * q, v);
                    tmem_ld_wait();
                    uint32_t o[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) {
                        const int r0 = 16 * q + j;
                        const float2 bb = *reinterpret_cast<const float2*>(br + r0);
                        float v0 = x[r0] + (__uint_as_float(v[j]) + bb.x), v1 = x[r0 + 1] + (__uint_as_float(v[j + 1]) + bb.y);
                        if (!valid) { v0 = 0.f; v1 = 0.f; }
                        x[r0] = v0; x[r0 + 1] = v1;
                        o[j >> 1] = pack_h2(v0, v1);
                        if (dump && valid) { p.xtOut[((size_t)l * B + b) * R + r0] = v0; p.xtOut[((size_t)l * B + b) * R + r0 + 1] = v1; }
                    }
                    if (l + 1 < L) {
                        const uint4 o0 = make_uint4(o[0], o[1], o[2], o[3]), o1 = make_uint4(o[4], o[5], o[6], o[7]);
                        *reinterpret_cast<uint4*>(t_xc + chunk_off(row, 2 * q)) = o0;
                        *reinterpret_cast<uint4*>(t_xc + chunk_off(row, 2 * q + 1)) = o1;
                    }
                }
                if (l + 1 < L) {
                    tc_fence_before_sync();
                    fence_proxy_async();
                    mbar_arrive(epi_done);                              // x_{l+1} ready
                    if (tid == 0) TRACE(0, 5);
                    store_history(ring_tile(t, l + 1));
                }
                // conditioning of the next layer (next sample when wrapping): issued here, consumed after the next MMA;
                // the rows of the layer after that are pulled into L2 now
                {
                    const bool wrap = (l + 1 == L);
                    const int tn = wrap ? t + 1 : t, ln = wrap ? 0 : l + 1;
                    if (tn < t_end) {
                        load_lh(tn, ln);
                        const bool wrap2 = (ln + 1 == L);
                        const int t2 = wrap2 ? tn + 1 : tn, l2 = wrap2 ? 0 : ln + 1;
                        if (t2 < t_end) { prefetch_l2(lh_ptr(t2, l2)); prefetch_l2(reinterpret_cast<const char*>(lh_ptr(t2, l2)) + 128); }
                    }
                }
                // ---------------- per-layer skip dump (last sample of a dumping launch only)
                if (dump && l + 1 < L) {
                    mbar_wait(skip_full, ph_skip); ph_skip ^= 1;
                    tc_fence_after_sync();
                    for (int c0 = 0; c0 < S; c0 += 16) {
                        uint32_t v[16];
                        tmem_ld16(DSKIP + lane_off + c0, v);
                        tmem_ld_wait();
                        if (valid)
                            for (int j = 0; j < 16; j++)
                                p.skipOut[((size_t)l * B + b) * S + c0 + j] = __uint_as_float(v[j]) + gbias[im.b_bskp + (size_t)l * S + c0 + j];
                    }
                    tc_fence_before_sync();
                }
            }

            // ---------------- relu(skip) -> GEMM input of the first output layer   (reference.cpp:88-90)
            mbar_wait(skip_full, ph_skip); ph_skip ^= 1;
            tc_fence_after_sync();
            if (tid == 0) TRACE(0, 6);
#pragma unroll 1
            for (int c0 = 0; c0 < S; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(DSKIP + lane_off + c0, v);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float v0 = fmaxf(__uint_as_float(v[j]) + s_bsk[c0 + j], 0.f), v1 = fmaxf(__uint_as_float(v[j + 1]) + s_bsk[c0 + j + 1], 0.f);
                    if (!valid) { v0 = 0.f; v1 = 0.f; }
                    o[j >> 1] = pack_h2(v0, v1);
                    if (dump && valid) { p.skipOut[((size_t)(L - 1) * B + b) * S + c0 + j] = v0; p.skipOut[((size_t)(L - 1) * B + b) * S + c0 + j + 1] = v1; }
                }
                unsigned char* kt = t_big + (size_t)(c0 >> 6) * TILE;
                const int q = (c0 & 63) >> 3;
                *reinterpret_cast<uint4*>(kt + chunk_off(row, q)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(kt + chunk_off(row, q + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            tc_fence_before_sync();
            fence_proxy_async();
            mbar_arrive(epi_done);                                      // relu(skip) tile ready
            if (tid == 0) TRACE(0, 7);

            // ---------------- Zs = relu(Wzs . skip + Bzs)   (reference.cpp:96-98)
            mbar_wait(out_full, ph_out); ph_out ^= 1;
            tc_fence_after_sync();
#pragma unroll 1
            for (int c0 = 0; c0 < A; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(DZS + lane_off + c0, v);
                tmem_ld_wait();
                uint32_t o[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float v0 = fmaxf(__uint_as_float(v[j]) + s_bzs[c0 + j], 0.f), v1 = fmaxf(__uint_as_float(v[j + 1]) + s_bzs[c0 + j + 1], 0.f);
                    if (!valid) { v0 = 0.f; v1 = 0.f; }
                    o[j >> 1] = pack_h2(v0, v1);
                    if (dump && valid) { p.Zs[(size_t)b * A + c0 + j] = v0; p.Zs[(size_t)b * A + c0 + j + 1] = v1; }
                }
                unsigned char* kt = t_big + (size_t)(c0 >> 6) * TILE;
                const int q = (c0 & 63) >> 3;
                *reinterpret_cast<uint4*>(kt + chunk_off(row, q)) = make_uint4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<uint4*>(kt + chunk_off(row, q + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
            }
            tc_fence_before_sync();
            fence_proxy_async();
            mbar_arrive(epi_done);                                      // relu(Zs) tile ready
            if (tid == 0) TRACE(0, 9);

            // ---------------- Za, softmax, categorical sample -- all inside this thread   (reference.cpp:100-121)
            mbar_wait(out_full, ph_out); ph_out ^= 1;
            tc_fence_after_sync();
            if (tid == 0) TRACE(0, 10);
            float mx = 0.f;                                             // matrix.cpp:171 starts the max at 0
#pragma unroll 1
            for (int c0 = 0; c0 < A; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(DZA + lane_off + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j++) mx = fmaxf(mx, __uint_as_float(v[j]) + s_bza[c0 + j]);
            }
            float csum[A / 16];
            float total = 0.f;
#pragma unroll
            for (int c = 0; c < A / 16; c++) {
                uint32_t v[16];
                tmem_ld16(DZA + lane_off + 16 * c, v);
                tmem_ld_wait();
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 16; j++) s += wn::exp2f_fast((__uint_as_float(v[j]) + s_bza[16 * c + j] - mx) * 1.4426950408889634f);
                csum[c] = s;
                total += s;
            }
            const float target = sel * total;
            int cb = A / 16 - 1;
            float base = 0.f;
            {
                float run = 0.f;
                bool found = false;
#pragma unroll
                for (int c = 0; c < A / 16; c++) {
                    if (!found && target < run + csum[c]) { cb = c; base = run; found = true; }
                    run += csum[c];
                }
                if (!found) base = run - csum[A / 16 - 1];
            }
            // third pass: every lane walks all chunks (tcgen05.ld is warp-collective: uniform address, no divergence
            // around it) and scans only inside its own chunk `cb`; the dump of Za / P rides along
            int y = A - 1;
            {
                const float inv = 1.f / total;
                bool found = false;
#pragma unroll 1
                for (int c = 0; c < A / 16; c++) {
                    uint32_t v[16];
                    tmem_ld16(DZA + lane_off + 16 * c, v);
                    tmem_ld_wait();
                    if (c == cb || dump) {
                        float run = base;
#pragma unroll
                        for (int j = 0; j < 16; j++) {
                            const float z = __uint_as_float(v[j]) + s_bza[16 * c + j];
                            const float e = wn::exp2f_fast((z - mx) * 1.4426950408889634f);
                            if (c == cb) {
                                run += e;
                                if (!found && target < run) { y = 16 * c + j; found = true; }
                            }
                            if (dump && valid) { p.Za[(size_t)b * A + 16 * c + j] = z; p.P[(size_t)b * A + 16 * c + j] = e * inv; }
                        }
                    }
                }
            }
            if (valid) {
                p.yOut[(size_t)b * p.N + t] = y;
                const int fb = p.forced ? p.forced[(size_t)b * p.N + t] : y;
                yp = yc;
                yc = fb;
            }
            if (tid == 0) TRACE(0, 11);
            // Dza is consumed: the x_0-ready arrival of the next sample (or kernel end) releases it
        }
        if (valid) { p.yPrev[b] = yp; p.yCur[b] = yc; }
        tc_fence_before_sync();
    }

    __syncthreads();
    if (warp == 4) tmem_dealloc<512>(tmem_base);
}

int pick_nstage(int S, int L)
{
    for (int n = 8; n >= 3; n--)
        if (tc_smem_bytes(S, L, n) <= 227 * 1024) return n;
    return 0;
}

}  // namespace

bool wn_tc_supported(int R_, int S, int A_, int L, int)
{
    return R_ == R && A_ == A && (S == 128 || S == 256) && pick_nstage(S, L) >= 3;
}

size_t wn_tc_image_bytes(int, int S, int, int L) { return tc_image(S, L).total; }

size_t wn_tc_ring_bytes(int S, int L, int maxDil, int B) { return (size_t)(maxDil + 1) * L * ((B + 127) / 128) * TILE; }

cudaError_t wn_tc_pack(void* image, const WnParams& p, cudaStream_t stream)
{
    const TcImage im = tc_image(p.S, p.L);
    cudaError_t e = cudaMemsetAsync(image, 0, im.total, stream);
    if (e != cudaSuccess) return e;
    tc_pack_kernel<<<296, 256, 0, stream>>>(p, static_cast<unsigned char*>(image), im);
    return cudaGetLastError();
}

cudaError_t wn_launch_tc(const WnParams& p, const void* tc_image_, cudaStream_t stream, WnLaunchInfo* info)
{
    const int nstage = pick_nstage(p.S, p.L);
    if (nstage < 3) return cudaErrorInvalidValue;
    const size_t smem = tc_smem_bytes(p.S, p.L, nstage);
    const int grid = (p.B + 127) / 128;
    cudaError_t e;
    if (p.S == 256) {
        e = cudaFuncSetAttribute(wn_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wn_tc_kernel<256><<<grid, NT, smem, stream>>>(p, static_cast<const unsigned char*>(tc_image_), nstage);
    } else {
        e = cudaFuncSetAttribute(wn_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        wn_tc_kernel<128><<<grid, NT, smem, stream>>>(p, static_cast<const unsigned char*>(tc_image_), nstage);
    }
    if (info) { info->kernel = 17; info->grid = grid; info->block = NT; info->smem_bytes = (int)smem; info->batch_per_cta = 128; info->cluster = 1; }
    return cudaGetLastError();
}
