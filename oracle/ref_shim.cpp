/*
 * ref_shim.cpp -- extern "C" door into the UNMODIFIED reference CPU model.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with
 * /root/reference/{nv_wavenet_reference.cpp,matrix.cpp} (read where they lie,
 * never copied) into oracle/_ref/libnvwn_ref.so.  The shim only forwards to the
 * reference's own classes: nvWavenetReference (nv_wavenet_reference.h:36-101)
 * and Matrix::randomize (matrix.cpp:38-56).
 *
 * ref_gen_test_inputs() replays the input construction of the reference's
 * integration test (nv_wavenet_test.cu:44-111, 217-219) call for call, so that
 * with the same srand() seed it yields the very weights / Lh / selectors the
 * reference test feeds to both of its implementations.
 */
#include "matrix.h"
#include "nv_wavenet_reference.h"

#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {

void* ref_create(int num_layers, int max_batch, int max_samples, int R, int S, int A, int max_dilation)
{
    return new nvWavenetReference(num_layers, max_batch, max_samples, R, S, A, max_dilation);
}
/* The reference destructor frees Matrix objects that themselves never free their
 * storage (matrix.h:31-56); we leak the same way rather than touch it. */
void ref_destroy(void* p) { delete (nvWavenetReference*)p; }

void ref_set_embeddings(void* p, float* prev, float* cur) { ((nvWavenetReference*)p)->setEmbeddings(prev, cur); }
void ref_set_layer_weights(void* p, int layer, float* Wprev, float* Wcur, float* Bh, float* Wres, float* Bres, float* Wskip, float* Bskip)
{
    ((nvWavenetReference*)p)->setLayerWeights(layer, Wprev, Wcur, Bh, Wres, Bres, Wskip, Bskip);
}
void ref_set_out_weights(void* p, float* Wzs, float* Bzs, float* Wza, float* Bza) { ((nvWavenetReference*)p)->setOutWeights(Wzs, Bzs, Wza, Bza); }
void ref_set_inputs(void* p, float* Lh, float* selectors) { ((nvWavenetReference*)p)->setInputs(Lh, selectors); }
void ref_get_xt_out(void* p, int layer, float* out) { ((nvWavenetReference*)p)->getXtOut(layer, out); }
void ref_get_skip_out(void* p, int layer, float* out) { ((nvWavenetReference*)p)->getSkipOut(layer, out); }
void ref_get_zs(void* p, float* out) { ((nvWavenetReference*)p)->getZs(out); }
void ref_get_za(void* p, float* out) { ((nvWavenetReference*)p)->getZa(out); }
void ref_get_p(void* p, float* out) { ((nvWavenetReference*)p)->getP(out); }
void ref_run(void* p, int num_samples, int batch_size, int* yOut) { ((nvWavenetReference*)p)->run(num_samples, batch_size, yOut); }

void ref_srand(unsigned seed) { srand(seed); }
int ref_rand(void) { return rand(); }

/* Matrix::randomize on a caller buffer (col-major rows x cols). */
void ref_randomize(float* dst, int rows, int cols, float mean, float scale)
{
    Matrix m(rows, cols, false);
    m.randomize(mean, scale);
    memcpy(dst, m.data(), sizeof(float) * rows * cols);
    free(m.data());
}

static void fill(float* dst, int rows, int cols, float mean, float scale)
{
    ref_randomize(dst, rows, cols, mean, scale);
}

/* createMatrix of the reference test (nv_wavenet_test.cu:36-42): scale = 0.5 / rows. */
static void create_matrix(float* dst, int r, int c)
{
    float mean = 0.0;
    float scale = 0.5 / r;
    fill(dst, r, c, mean, scale);
}

/*
 * Same rand() consumption order as runTest<>() (nv_wavenet_test.cu:44-220).
 * Output buffers (caller allocated):
 *   selectors [N][B]  embPrev/embCur [A][R]
 *   Wprev/Wcur [L][2R*R]  Bh [L][2R]  Wres [L][R*R]  Bres [L][R]  Wskip [L][S*R]  Bskip [L][S]
 *   Wzs [A*S] Bzs [A] Wza [A*A] Bza [A]   Lh [N][L][B][2R]
 */
void ref_gen_test_inputs(int R, int S, int A, int L, int B, int N,
                         float* selectors, float* embPrev, float* embCur,
                         float* Wprev, float* Wcur, float* Bh, float* Wres, float* Bres, float* Wskip, float* Bskip,
                         float* Wzs, float* Bzs, float* Wza, float* Bza, float* Lh)
{
    float mean = 0.0;
    float scale = 0.5 / R;
    for (int b = 0; b < B; b++) { (void)(rand() % A); (void)(rand() % A); }      /* :54-57 */
    fill(selectors, B, N, 0.5, 1.0);                                             /* :60-61 */
    fill(embPrev, R, A, mean, scale);                                            /* :66-67 */
    fill(embCur, R, A, mean, scale);
    std::vector<float> scratch((size_t)S * B > (size_t)R * B ? (size_t)S * B : (size_t)R * B);
    for (int l = 0; l < L; l++) {                                                /* :84-96 */
        create_matrix(Wprev + (size_t)l * 2 * R * R, 2 * R, R);
        create_matrix(Wcur + (size_t)l * 2 * R * R, 2 * R, R);
        create_matrix(Bh + (size_t)l * 2 * R, 2 * R, 1);
        create_matrix(Wres + (size_t)l * R * R, R, R);
        create_matrix(Bres + (size_t)l * R, R, 1);
        create_matrix(Wskip + (size_t)l * S * R, S, R);
        create_matrix(Bskip + (size_t)l * S, S, 1);
        create_matrix(scratch.data(), S, B);                                     /* skipOut[l] */
    }
    for (int s = 0; s < N; s++)                                                  /* :98-102 */
        for (int l = 0; l < L + 1; l++) create_matrix(scratch.data(), R, B);
    fill(Wzs, A, S, mean, scale);                                                /* :104-111 */
    fill(Bzs, A, 1, mean, scale);
    fill(Wza, A, A, mean, scale);
    fill(Bza, A, 1, mean, scale);
    fill(Lh, 2 * R, N * L * B, mean, scale);                                     /* :217-219 */
}

}  /* extern "C" */
