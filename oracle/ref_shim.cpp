// TEST INFRASTRUCTURE — not product code.
//
// oracle/_ref/libsnn_ref.so: a thin extern "C" shim that #includes the reference's OWN
// CPU operator code where it lies under /root/reference (nothing is copied into this
// repo) so that the oracle restatement in snn_oracle.cpp can be pinned against the real
// thing, and so that bench.py can time the reference's Dense/softmax tail.
//
// What the reference offers on CPU: exactly one header,
//   core/src/ic2/cpulayer.h  — snn::dp::CPUCommonUtil<T>: Dense (Eigen W*x+b), activations, softmax
// plus the unit-test PRNG
//   demo/common/prng.h       — lagged-Fibonacci generator seeded with 7767517 in every op test.
// Everything else (Conv/Depthwise/Pool/BN) exists only as GLSL and cannot run here.
//
// Build recipe: oracle/Makefile target `_ref` (g++ on this one file, include paths into
// /root/reference; no reference build system is run).
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include "snn/image.h"
#include "ic2/cpulayer.h"

extern "C" {
#include "prng.h"
}

// The four symbols cpulayer.h needs from libsnn_core (which cannot be built here: OpenCV/GL/Vulkan absent).
namespace snn {
bool isLoggable(int, int&, const char*) { return false; }
void log(const char*, int, const char*, int, int, const char*, ...) {}
[[noreturn]] void rip() { std::abort(); }
float convertToHighPrecision(uint16_t) { return 0.f; }
} // namespace snn

extern "C" {

// Dense + activation exactly as DenseLayer::computeImageTexture drives CPUCommonUtil
// (core/src/ic2/denselayer.cpp:27-54): weights arrive as the parser's vector<vector<float>>
// [numInputUnits][numOutputUnits] filled from the flat JSON kernel (modelparser.cpp:524-535).
// flat_kernel has n_in*n_out floats in JSON order.
int ref_dense(const float* x, int n_in, const float* flat_kernel, const float* bias, int n_out, const char* activation, float alpha, float* y) {
    std::vector<std::vector<float>> w(n_in, std::vector<float>(n_out));
    size_t e = 0;
    for (int i = 0; i < n_in; ++i)
        for (int j = 0; j < n_out; ++j) w[i][j] = flat_kernel[e++];
    std::vector<float> b(bias, bias + n_out);
    snn::dp::CPUCommonUtil<float> util(activation, alpha, true);
    util.inputMat = std::vector<std::vector<float>> {std::vector<float>(x, x + n_in)};
    std::pair<std::vector<std::vector<float>>, std::vector<float>> tm(w, b);
    util.run(tm);
    auto out = util.getOutputs();
    if (out.size() != 1 || (int) out[0].size() != n_out) return -1;
    for (int j = 0; j < n_out; ++j) y[j] = out[0][j];
    return 0;
}

// Activation only (identity transform is taken when the weight list is empty AND inputMat is set?
// no: cpulayer.h:151 takes the identity branch when first.empty() — used by FlattenLayer).
int ref_activation(const float* x, int n, const char* activation, float alpha, float* y) {
    snn::dp::CPUCommonUtil<float> util(activation, alpha, true);
    util.inputMat = std::vector<std::vector<float>> {std::vector<float>(x, x + n)};
    std::pair<std::vector<std::vector<float>>, std::vector<float>> tm;
    util.run(tm);
    auto out = util.getOutputs();
    // identity path emits outputs.colwise() of an n-vector => one row of n values
    size_t k = 0;
    for (auto& row : out)
        for (auto v : row) {
            if ((int) k >= n) return -1;
            y[k++] = v;
        }
    return (int) k == n ? 0 : -1;
}

// demo/common/prng.h driven the way demo/common/testutil.cpp:40-47 does (RandomFloat).
static struct prng_rand_t g_state;
void ref_srand(uint64_t seed) { prng_srand(seed, &g_state); }
uint64_t ref_rand_u64() { return prng_rand(&g_state); }
float ref_random_float(float a, float b) {
    float random = ((float) prng_rand(&g_state)) / (float) uint64_t(-1);
    float diff   = b - a;
    float rd     = random * diff;
    return a + rd;
}
}
