/*
 * ref_gpu_harness.cu -- runs the reference's OWN CUDA kernels (unmodified headers, read in place from
 * /root/reference) on given inputs.  TEST / BASELINE INFRASTRUCTURE ONLY: gives (a) the GPU baseline the new
 * kernels have to beat on the same box (reference PERSISTENT kernel rebuilt for sm_100a) and (b) a GPU oracle
 * for parity (sampled indices in fp32, logits in fp16).
 *
 * Build (oracle/Makefile, target refgpu): nvcc -arch=sm_100a --use_fast_math -I/root/reference -DmemoryType=type
 *   (-DmemoryType=type: cudaPointerAttributes::memoryType was renamed `type` in CUDA >= 11,
 *    nv_wavenet_conversions.cuh:41; the macro renames the one token without touching the source).
 *
 * usage: ref_gpu_harness <input.bin> <output.bin>
 *   input : int32 header {precision(16|32), R, S, A, L, maxDil, B, N, mode, tanhEmbed, chunk, reps} then fp32 arrays
 *           embPrev[A*R] embCur[A*R] {Wprev Wcur Bh Wres Bres Wskip Bskip}xL Wzs Bzs Wza Bza Lh[N*L*B*2R] sel[N*B]
 *   output: float elapsed_ms (best of reps), int32 yOut[B*N], float Za[B*A], float P[B*A]
 */
#include "nv_wavenet.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <vector>

template <typename TW, typename TD, int R, int S, int A>
static int run(const int* h, const float* f, FILE* out)
{
    const int L = h[4], maxDil = h[5], B = h[6], N = h[7], mode = h[8], tanhEmbed = h[9], chunk = h[10], reps = h[11];
    nvWavenetInfer<TW, TD, R, S, A> infer(L, maxDil, B, N, mode, tanhEmbed != 0);
    const float* p = f;
    auto take = [&](size_t n) { const float* q = p; p += n; return const_cast<float*>(q); };
    float* embPrev = take((size_t)A * R); float* embCur = take((size_t)A * R);
    infer.setEmbeddings(embPrev, embCur);
    for (int l = 0; l < L; l++) {
        float* Wprev = take(2 * R * R); float* Wcur = take(2 * R * R); float* Bh = take(2 * R);
        float* Wres = take(R * R); float* Bres = take(R); float* Wskip = take(S * R); float* Bskip = take(S);
        infer.setLayerWeights(l, Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip);
    }
    float* Wzs = take((size_t)A * S); float* Bzs = take(A); float* Wza = take((size_t)A * A); float* Bza = take(A);
    infer.setOutWeights(Wzs, Bzs, Wza, Bza);
    float* Lh = take((size_t)N * L * B * 2 * R); float* sel = take((size_t)N * B);
    const int bspb = (B % 4 == 0) ? 4 : (B % 2 == 0) ? 2 : 1;
    int* yOut;
    gpuErrChk(cudaMallocHost(&yOut, (size_t)N * B * sizeof(int)));
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        infer.setInputs(Lh, sel);
        gpuErrChk(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        bool ok = infer.run_chunks(chunk, [](int*, int, int) {}, N, B, yOut, bspb);   // as nv_wavenet_perf.cu:75
        cudaEventRecord(e1);
        gpuErrChk(cudaEventSynchronize(e1));
        gpuErrChk(cudaDeviceSynchronize());
        if (!ok) return 3;
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    std::vector<float> Za((size_t)B * A), P((size_t)B * A);
    infer.getZa(Za.data());
    infer.getP(P.data());
    fwrite(&best, sizeof(float), 1, out);
    fwrite(yOut, sizeof(int), (size_t)N * B, out);
    fwrite(Za.data(), sizeof(float), Za.size(), out);
    fwrite(P.data(), sizeof(float), P.size(), out);
    printf("{\"ref_gpu_ms\": %f, \"khz_per_utterance\": %f, \"samples_per_s\": %f}\n", best, N / best, (double)N * B / (best * 1e-3));
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
    FILE* in = fopen(argv[1], "rb");
    if (!in) { perror("input"); return 2; }
    int h[12];
    if (fread(h, sizeof(int), 12, in) != 12) return 2;
    fseek(in, 0, SEEK_END);
    const long bytes = ftell(in) - 12 * (long)sizeof(int);
    fseek(in, 12 * sizeof(int), SEEK_SET);
    std::vector<float> f(bytes / sizeof(float));
    if (fread(f.data(), sizeof(float), f.size(), in) != f.size()) return 2;
    fclose(in);
    FILE* out = fopen(argv[2], "wb");
    if (!out) { perror("output"); return 2; }
    int rc = 4;
    const int prec = h[0], R = h[1], S = h[2], A = h[3];
    if (R == 64 && S == 256 && A == 256) rc = (prec == 16) ? run<half2, half, 64, 256, 256>(h, f.data(), out) : run<float, float, 64, 256, 256>(h, f.data(), out);
    else if (R == 64 && S == 128 && A == 256 && prec == 16) rc = run<half2, half, 64, 128, 256>(h, f.data(), out);      // BASELINE.json configs[1] (C2)
    else if (R == 128 && S == 256 && A == 256 && prec == 32) rc = run<float, float, 128, 256, 256>(h, f.data(), out);  // BASELINE.json configs[3] (C4)
    else fprintf(stderr, "unsupported shape (built for R64/S256/A256 fp16+fp32, R64/S128/A256 fp16, R128/S256/A256 fp32)\n");
    fclose(out);
    return rc;
}
