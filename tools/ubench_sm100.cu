// ubench_sm100.cu -- micro-benchmarks that fix the design parameters of the latency-mode kernel (wn_lat_kernel.cu).
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench tools/ubench_sm100.cu
// Prints one JSON object per line.  Every loop is bounded; nothing can hang.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void hmma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---------------------------------------------------------------- 1. HMMA latency / throughput
// NACC independent accumulators per warp, ITER rounds; reports cycles per HMMA per warp
template <int NACC>
__global__ void k_hmma(int iters, long long* out, float* sink)
{
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x38003800u, 0x38003800u};
    a[0] += threadIdx.x;
    float d[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) hmma(d[i], a, b);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

// ---------------------------------------------------------------- 2. HMMA fed from shared memory (B fragments by LDS.128, 2 HMMA per load)
// each warp owns its own 16 KB slice of a weight image in smem; per round: LDS.128 -> 2 HMMA
__global__ void k_hmma_lds(int iters, long long* out, float* sink)
{
    extern __shared__ __align__(16) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    const int per = (128 * 1024) / nw;                    // bytes per warp
    for (int i = threadIdx.x; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x38003800u;
    __syncthreads();
    uint32_t a[4] = {0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u};
    float d[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
    const unsigned char* base = sm + (size_t)warp * per + lane * 16;
    const int nld = per / 512;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll 8
        for (int j = 0; j < nld; j++) {
            const uint4 w = *reinterpret_cast<const uint4*>(base + (size_t)j * 512);
            const uint32_t b0[2] = {w.x, w.y}, b1[2] = {w.z, w.w};
            hmma(d[(2 * j) & 7], a, b0);
            hmma(d[(2 * j + 1) & 7], a, b1);
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
    if (s == 12345.f) *sink = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = (long long)nld * 2 * iters; }
}

// ---------------------------------------------------------------- 3. L2 -> SM streaming of one CTA
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity)
{
    for (uint32_t s = 0; s < (1u << 22); s++) if (mbar_try(bar, parity)) return true;
    return false;
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one producer thread streams `total` bytes in `chunk`-byte pieces through NST stages; consumer = thread 32 releases at once
__global__ void k_stream_tma(const unsigned char* src, size_t span, int chunk, int nst, int nchunks, long long* out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    uint64_t* full = reinterpret_cast<uint64_t*>(sm + (size_t)nst * chunk);
    uint64_t* empty = full + nst;
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const unsigned char* base = src + (size_t)blockIdx.x * 4096;       // CTAs read the same (L2-resident) data
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        int st = 0; uint32_t ph = 1; size_t off = 0;
        for (int c = 0; c < nchunks; c++) {
            if (!mbar_wait(&empty[st], ph)) break;
            mbar_expect(&full[st], chunk);
            tma_load_1d(sm + (size_t)st * chunk, base + off, chunk, &full[st]);
            off += chunk; if (off + chunk + 4096 * gridDim.x > span) off = 0;
            if (++st == nst) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        int st = 0; uint32_t ph = 0;
        for (int c = 0; c < nchunks; c++) {
            if (!mbar_wait(&full[st], ph)) break;
            mbar_arrive(&empty[st]);
            if (++st == nst) { st = 0; ph ^= 1; }
        }
        out[blockIdx.x] = clock64() - t0;
    }
}
// all threads of the CTA stream with LDG.128 (L1 bypass), `unroll` loads in flight per thread
__global__ void k_stream_ldg(const uint4* src, size_t n16, int rounds, long long* out, uint32_t* sink)
{
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        for (size_t i = threadIdx.x; i + 7 * blockDim.x < n16; i += 8 * blockDim.x) {
            uint4 v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[k].x), "=r"(v[k].y), "=r"(v[k].z), "=r"(v[k].w) : "l"(src + i + (size_t)k * blockDim.x));
#pragma unroll
            for (int k = 0; k < 8; k++) acc ^= v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
        }
    }
    const long long t1 = clock64();
    if (acc == 0x12345u) *sink = acc;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- 4. bar.sync, and the STS.128 -> bar -> LDS.128 exchange
__global__ void k_bar(int iters, int nthreads_bar, long long* out)
{
    __syncthreads();
    const long long t0 = clock64();
    if ((int)threadIdx.x < nthreads_bar)
        for (int i = 0; i < iters; i++) asm volatile("bar.sync 1, %0;" ::"r"(nthreads_bar) : "memory");
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
__global__ void k_exchange(int iters, long long* out, uint32_t* sink)
{
    __shared__ __align__(16) uint4 buf[2][256];
    const int n = blockDim.x;
    uint4 v = make_uint4(threadIdx.x, 1, 2, 3);
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        buf[i & 1][threadIdx.x] = v;
        asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory");
        const uint4 a = buf[i & 1][(threadIdx.x + 32) % n], b = buf[i & 1][(threadIdx.x + 64) % n];
        v.x = a.x + b.y; v.y = a.y ^ b.x; v.z = a.z + b.w; v.w = a.w + b.z;
    }
    const long long t1 = clock64();
    if (v.x == 0x1234567u) *sink = v.x;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

// ---------------------------------------------------------------- 5. DSMEM ping-pong inside a 2-CTA cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v)
{
    asm volatile("st.shared::cluster.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t addr)
{
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_cluster(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// 128 threads of CTA r write 16 B each (2 KB) into the peer's buffer and arrive on the peer's mbarrier (count 128);
// the peer waits, then answers the same way.  Reports cycles per one-way hop.
__global__ void __cluster_dims__(2, 1, 1) k_dsmem(int iters, long long* out)
{
    __shared__ __align__(16) uint4 buf[128];
    __shared__ uint64_t bar;
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t rank = cl.block_rank(), peer = rank ^ 1;
    if (threadIdx.x == 0) { mbar_init(&bar, 128); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    cl.sync();
    const uint32_t rbuf = mapa(smem_u32(&buf[threadIdx.x]), peer), rbar = mapa(smem_u32(&bar), peer);
    uint4 v = make_uint4(threadIdx.x, rank, 0, 0);
    uint32_t ph = 0;
    bool ok = true;
    const long long t0 = clock64();
    for (int i = 0; i < iters && ok; i++) {
        if ((i & 1) == (int)rank) {
            st_cluster_v4(rbuf, v);
            mbar_arrive_cluster(rbar);
        } else {
            ok = false;
            for (uint32_t s = 0; s < (1u << 22); s++) if (mbar_try_cluster(&bar, ph)) { ok = true; break; }
            ph ^= 1;
            v = buf[threadIdx.x]; v.z++;
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && rank == 0) { out[0] = t1 - t0; out[1] = ok ? 1 : 0; out[2] = v.z; }
    cl.sync();
}

// ---------------------------------------------------------------- 6. MUFU tanh.f16x2 / ex2 throughput per warp
__global__ void k_mufu(int iters, long long* out, uint32_t* sink)
{
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 8; i++) h[i] = 0x30003000u + threadIdx.x + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= h[i];
    if (s == 0x1234567u) *sink = s;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}

int main()
{
    long long* d_out; float* d_sink;
    CK(cudaMalloc(&d_out, 1024 * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 64));
    long long h[1024];
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, prop.multiProcessorCount, clk);
    auto fetch = [&]() { CK(cudaDeviceSynchronize()); CK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost)); };
    const int IT = 2000;
    // 1. HMMA
    for (int warps : {1, 4, 8, 16}) {
#define RUN_HMMA(NACC) do { k_hmma<NACC><<<1, warps * 32>>>(IT, d_out, d_sink); k_hmma<NACC><<<1, warps * 32>>>(IT, d_out, d_sink); fetch(); \
        printf("{\"bench\": \"hmma_m16n8k16_f32\", \"warps\": %d, \"indep_acc\": %d, \"cycles_per_hmma_per_warp\": %.2f, \"cycles_per_hmma_per_smsp\": %.2f}\n", warps, NACC, \
               (double)h[0] / (IT * NACC), (double)h[0] / (IT * NACC) / ((warps + 3) / 4)); } while (0)
        RUN_HMMA(1); RUN_HMMA(2); RUN_HMMA(4); RUN_HMMA(8);
    }
    // 2. HMMA fed by LDS.128
    CK(cudaFuncSetAttribute(k_hmma_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    for (int warps : {4, 8, 16}) {
        k_hmma_lds<<<1, warps * 32, 128 * 1024>>>(50, d_out, d_sink); fetch();
        printf("{\"bench\": \"hmma_fed_by_lds128\", \"warps\": %d, \"cycles_per_hmma_per_smsp\": %.2f, \"smem_bytes_per_clk\": %.1f}\n", warps,
               (double)h[0] / h[1] / ((warps + 3) / 4), (double)h[1] * 256.0 * warps / h[0]);
    }
    // 3. streaming
    {
        const size_t span = 4u << 20;
        unsigned char* d_src; CK(cudaMalloc(&d_src, span)); CK(cudaMemset(d_src, 1, span));
        CK(cudaFuncSetAttribute(k_stream_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int ctas : {1, 4, 148}) for (int chunk : {4096, 16384, 32768, 65536}) for (int nst : {2, 3}) {
            const int nch = (int)((16u << 20) / chunk);
            k_stream_tma<<<ctas, 64, (size_t)nst * chunk + 256>>>(d_src, span, chunk, nst, nch, d_out); fetch();
            long long mx = 0; for (int i = 0; i < ctas; i++) if (h[i] > mx) mx = h[i];
            printf("{\"bench\": \"l2_to_smem_tma\", \"ctas\": %d, \"chunk\": %d, \"stages\": %d, \"bytes_per_clk_per_cta\": %.2f}\n", ctas, chunk, nst, (double)nch * chunk / mx);
        }
        for (int ctas : {1, 4, 148}) for (int thr : {256, 512}) {
            const size_t n16 = (2u << 20) / 16;
            uint32_t* snk = reinterpret_cast<uint32_t*>(d_sink);
            k_stream_ldg<<<ctas, thr>>>(reinterpret_cast<const uint4*>(d_src), n16, 8, d_out, snk); fetch();
            long long mx = 0; for (int i = 0; i < ctas; i++) if (h[i] > mx) mx = h[i];
            printf("{\"bench\": \"l2_to_reg_ldg128\", \"ctas\": %d, \"threads\": %d, \"bytes_per_clk_per_cta\": %.2f}\n", ctas, thr, 8.0 * (2u << 20) / mx);
        }
        cudaFree(d_src);
    }
    // 4. barriers
    for (int n : {64, 128, 256}) {
        k_bar<<<1, 256>>>(IT, n, d_out); fetch();
        printf("{\"bench\": \"bar_sync\", \"threads\": %d, \"cycles\": %.1f}\n", n, (double)h[0] / IT);
        k_exchange<<<1, n>>>(IT, d_out, reinterpret_cast<uint32_t*>(d_sink)); fetch();
        printf("{\"bench\": \"sts128_bar_lds128_roundtrip\", \"threads\": %d, \"cycles\": %.1f}\n", n, (double)h[0] / IT);
    }
    // 5. DSMEM
    {
        k_dsmem<<<2, 128>>>(IT, d_out); fetch();
        printf("{\"bench\": \"dsmem_2KB_hop_plus_mbarrier\", \"ok\": %lld, \"cycles_per_hop\": %.1f}\n", h[1], (double)h[0] / IT);
    }
    // 6. MUFU
    for (int warps : {1, 4, 8}) {
        k_mufu<<<1, warps * 32>>>(IT, d_out, reinterpret_cast<uint32_t*>(d_sink)); fetch();
        printf("{\"bench\": \"tanh_f16x2\", \"warps\": %d, \"cycles_per_instr_per_warp\": %.2f}\n", warps, (double)h[0] / (IT * 8));
    }
    return 0;
}
