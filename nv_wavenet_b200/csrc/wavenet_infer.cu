// wavenet_infer.cu -- the reference's C-ABI (include/wavenet_infer.h) on top of the B200 engine.
// Replaces pytorch/wavenet_infer.cu:40-149 of the reference; semantics documented in the header.
#include "../../include/wavenet_infer.h"
#include "../../include/nvwn_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

namespace {
// must match the wavenet channels (pytorch/wavenet_infer.cu:34-37)
const int A = 256;
const int R = 64;
const int S = 256;

// error convention of the reference: print "GPUassert: ..." and exit (nv_wavenet_util.cuh:34-40)
void check(int rc, const char* what)
{
    if (rc != 0) {
        fprintf(stderr, "GPUassert: %s (%s) %s %d\n", nvwn_last_error(), what, __FILE__, rc);
        exit(rc > 0 ? rc : 1);
    }
}
}  // namespace

extern "C" {

void wavenet_infer(int sample_count, int batch_size, float* embedding_prev, float* embedding_curr,
                   int num_layers, int max_dilation,
                   float** in_layer_weights_prev, float** in_layer_weights_curr, float** in_layer_biases,
                   float** res_layer_weights, float** res_layer_biases,
                   float** skip_layer_weights, float** skip_layer_biases,
                   float* conv_out_weight, float* conv_end_weight, int use_embed_tanh,
                   float* cond_input, int implementation, int* samples)
{
    if (!samples) { fprintf(stderr, "wavenet_infer: samples must not be NULL\n"); abort(); }   // assert(samples), wavenet_infer.cu:142
    int dtype = NVWN_FP32;
    if (const char* env = getenv("NVWN_PRECISION")) {
        if (!strcmp(env, "fp16")) dtype = NVWN_FP16;
    }
    nvwn_engine* e = nullptr;
    check(nvwn_create(&e, dtype, R, S, A, num_layers, max_dilation, batch_size, sample_count,
                      (implementation >= 0 && implementation <= 4) ? NVWN_KERNEL_AUTO : implementation, use_embed_tanh),
          "create");
    check(nvwn_set_embeddings(e, embedding_prev, embedding_curr), "setEmbeddings");
    for (int l = 0; l < num_layers; l++) {
        check(nvwn_set_layer_weights(e, l, in_layer_weights_prev[l], in_layer_weights_curr[l], in_layer_biases[l],
                                     res_layer_weights[l], res_layer_biases[l], skip_layer_weights[l], skip_layer_biases[l]),
              "setLayerWeights");
    }
    // "We didn't use biases on our outputs" (wavenet_infer.cu:75-82)
    std::vector<float> zero_bias(A, 0.f);
    check(nvwn_set_out_weights(e, conv_out_weight, zero_bias.data(), conv_end_weight, zero_bias.data()), "setOutWeights");

    // Matrix outputSelectors(batch_size, sample_count); outputSelectors.randomize(0.5, 1.0)
    // (wavenet_infer.cu:92-93, matrix.cpp:38-56): rows = batch visited outermost, two rand() per element,
    // column-major storage => selectors[sample * batch_size + b].
    std::vector<float> selectors((size_t)sample_count * batch_size);
    for (int b = 0; b < batch_size; b++) {
        for (int s = 0; s < sample_count; s++) {
            (void)(rand() % 100);                                   // sparsity draw (sparsity = 0)
            float r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
            r -= 0.5;
            r = r * 1.0f + 0.5f;
            selectors[(size_t)s * batch_size + b] = r;
        }
    }
    check(nvwn_set_inputs(e, cond_input, selectors.data()), "setInputs");
    check(nvwn_run(e, sample_count, batch_size, samples, /*dumpActivations=*/1, nullptr), "run");
    cudaError_t ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        fprintf(stderr, "GPUassert: %s %s %d\n", cudaGetErrorString(ce), __FILE__, __LINE__);
        exit(ce);
    }
    nvwn_destroy(e);
}

int get_R(void) { return R; }
int get_S(void) { return S; }
int get_A(void) { return A; }

}  // extern "C"
