/*
 * nv_wavenet.hpp -- C++ facade with the class surface of the reference's
 *   template <typename T_weight, typename T_data, int R=64, int S=128, int A=256> class nvWavenetInfer
 * (nv_wavenet.cuh:220-640) forwarding to the C-ABI of libwavenet_infer.so (include/nvwn_b200.h).
 *
 * The reference class is header-only and instantiates its kernels in the caller's translation unit;
 * this facade is plain host C++ (no nvcc needed): the sm_100a kernels are compiled once inside the
 * library for every supported (precision, R, S) and selected at run time.
 *
 *   T_weight / T_data : float / float  -> bit-exact fp32 path
 *                       half2 / half   -> fp16 tensor-core path
 *   Implementation    : the enum values of the reference are accepted; all map to the one kernel family.
 *   Errors            : CUDA / argument errors print "GPUassert: ..." and exit(code) like gpuErrChk
 *                       (nv_wavenet_util.cuh:34-40); launch failures make run*() return false.
 */
#ifndef NV_WAVENET_HPP
#define NV_WAVENET_HPP

#include <cuda_fp16.h>
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "nvwn_b200.h"

namespace nvwn_detail {
template <typename T_data> struct dtype_of;
template <> struct dtype_of<float> { enum { value = NVWN_FP32 }; };
template <> struct dtype_of<half> { enum { value = NVWN_FP16 }; };
inline void check(int rc, const char* file, int line)
{
    if (rc != 0) {
        fprintf(stderr, "GPUassert: %s %s %d\n", nvwn_last_error(), file, line);
        exit(rc > 0 ? rc : 1);
    }
}
}  // namespace nvwn_detail
#define NVWN_CHK(x) nvwn_detail::check((x), __FILE__, __LINE__)

template <typename T_weight, typename T_data, int R = 64, int S = 128, int A = 256>
class nvWavenetInfer {
public:
    enum Implementation { AUTO = 0, SINGLE_BLOCK, DUAL_BLOCK, PERSISTENT, MANYBLOCK_NONPERSISTENT };

protected:
    nvwn_engine* m_engine;
    int m_maxBatch, m_maxSamples, m_num_samples_per_chunk;

public:
    nvWavenetInfer(int numLayers, int maxDilation, int batchSize, int numSamples, int impl = 0, bool tanhEmbed = true)
        : m_engine(NULL), m_maxBatch(batchSize), m_maxSamples(numSamples), m_num_samples_per_chunk(0)
    {
        NVWN_CHK(nvwn_create(&m_engine, nvwn_detail::dtype_of<T_data>::value, R, S, A, numLayers, maxDilation,
                             batchSize, numSamples, impl, tanhEmbed ? 1 : 0));
    }
    virtual ~nvWavenetInfer() { nvwn_destroy(m_engine); }

    virtual void setEmbeddings(float* embedPrev, float* embedCur) { NVWN_CHK(nvwn_set_embeddings(m_engine, embedPrev, embedCur)); }
    virtual void setLayerWeights(int layer, float* Wprev, float* Wcur, float* Bh, float* Wres, float* Bres, float* Wskip, float* Bskip)
    {
        NVWN_CHK(nvwn_set_layer_weights(m_engine, layer, Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip));
    }
    virtual void setOutWeights(float* Wzs, float* Bzs, float* Wza, float* Bza) { NVWN_CHK(nvwn_set_out_weights(m_engine, Wzs, Bzs, Wza, Bza)); }
    void setInputs(float* Lh, float* outputSelectors) { NVWN_CHK(nvwn_set_inputs(m_engine, Lh, outputSelectors)); }

    void getXtOut(int layer, float* hXt) { NVWN_CHK(nvwn_get_xt_out(m_engine, layer, hXt)); }
    void getSkipOut(int layer, float* hSkipOut) { NVWN_CHK(nvwn_get_skip_out(m_engine, layer, hSkipOut)); }
    void getZs(float* hZs) { NVWN_CHK(nvwn_get_zs(m_engine, hZs)); }
    void getZa(float* hZa) { NVWN_CHK(nvwn_get_za(m_engine, hZa)); }
    void getP(float* hP) { NVWN_CHK(nvwn_get_p(m_engine, hP)); }
    void getYOut(int* yOut, int offset, int size, cudaStream_t stream = 0) { NVWN_CHK(nvwn_get_yout(m_engine, yOut, offset, size, stream)); }
    /* Not in the reference class: mu-law decoded audio of yOut[b][offset .. offset+size) for every b, on the device
     * (replaces the host post-processing of pytorch/nv_wavenet_inference.py:55-60); see nvwn_get_audio. */
    void getAudio(float* audio_f32, short* audio_i16, int offset, int size, bool saturate = false, cudaStream_t stream = 0)
    {
        NVWN_CHK(nvwn_get_audio(m_engine, audio_f32, audio_i16, offset, size, saturate ? 1 : 0, stream));
    }

    /* Chunked generation with overlapped device->host copies of finished chunks (nv_wavenet.cuh:445-497).
     * consume(yOut, initSample, count) is called per chunk once its copy has landed. */
    template <class Callback>
    bool run_chunks(int num_samples_per_chunk, Callback consume, int num_samples, int batch_size, int* yOut = NULL,
                    int batch_size_per_block = 1, bool dumpActivations = false, cudaStream_t stream = 0)
    {
        bool result = true;
        cudaStream_t stream_compute = stream, stream_copy;
        if (!stream) cudaStreamCreate(&stream_compute);
        cudaStreamCreate(&stream_copy);
        const int num_chunks = (num_samples + num_samples_per_chunk - 1) / num_samples_per_chunk;
        std::vector<cudaEvent_t> ev_compute(num_chunks), ev_copy(num_chunks);
        for (int j = 0; j < num_chunks; j++) {
            cudaEventCreateWithFlags(&ev_compute[j], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&ev_copy[j], cudaEventDisableTiming);
        }
        for (int j = 0; j < num_chunks; j++) {
            const int init = j * num_samples_per_chunk;
            m_num_samples_per_chunk = (j == num_chunks - 1) ? num_samples - init : num_samples_per_chunk;
            result = result && run_partial(init, num_samples, batch_size, NULL, batch_size_per_block, true, stream_compute);
            cudaEventRecord(ev_compute[j], stream_compute);
            cudaStreamWaitEvent(stream_copy, ev_compute[j], 0);
            if (yOut != NULL) getYOut(yOut, init, m_num_samples_per_chunk, stream_copy);
            cudaEventRecord(ev_copy[j], stream_copy);
        }
        for (int j = 0; j < num_chunks; j++) {
            const int init = j * num_samples_per_chunk;
            const int n = (j == num_chunks - 1) ? num_samples - init : num_samples_per_chunk;
            cudaEventSynchronize(ev_copy[j]);
            consume(yOut, init, n);
        }
        m_num_samples_per_chunk = 0;
        for (int j = 0; j < num_chunks; j++) { cudaEventDestroy(ev_compute[j]); cudaEventDestroy(ev_copy[j]); }
        if (stream != stream_compute) cudaStreamDestroy(stream_compute);
        cudaStreamDestroy(stream_copy);
        return result;
    }

    bool run_partial(int init_sample, int num_samples, int batch_size, int* yOut = NULL, int batch_size_per_block = 1,
                     bool dumpActivations = false, cudaStream_t stream = 0)
    {
        (void)batch_size_per_block;      /* the batch tile per CTA is chosen by the library */
        const int count = m_num_samples_per_chunk ? m_num_samples_per_chunk : num_samples;
        const int rc = nvwn_run_partial(m_engine, init_sample, count, num_samples, batch_size, yOut, dumpActivations ? 1 : 0, stream);
        if (rc != 0) fprintf(stderr, "GPUassert: %s %s %d\n", nvwn_last_error(), __FILE__, __LINE__);
        return rc == 0;
    }

    bool run(int num_samples, int batch_size, int* yOut = NULL, int batch_size_per_block = 1, bool dumpActivations = false,
             cudaStream_t stream = 0)
    {
        m_num_samples_per_chunk = 0;
        return run_partial(0, num_samples, batch_size, yOut, batch_size_per_block, dumpActivations, stream);
    }
};

#endif
