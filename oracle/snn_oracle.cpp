// TEST INFRASTRUCTURE — the parity ORACLE. Not product code: only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library. The product path
// (shadernn_b200/csrc) never links or calls it.
//
// A CPU restatement, in plain C++ (fp32, NHWC, batch-aware), of the operator semantics of
// inferenceengine/shadernn. The reference has NO CPU implementation of Conv/Depthwise/Pool/BN
//  — its only executable statement of that arithmetic is GLSL — so each function
// below restates the Vulkan compute shader + the host-side .cpp that parameterises it, and cites
// them (paths relative to /root/reference). Dense/softmax/activations restate core/src/ic2/cpulayer.h,
// which IS compilable here and against which this file is pinned (oracle/_ref, tests/test_oracle_ref.py).
//
// PARITY PINNING: Dense/softmax/activation/PRNG are pinned against the compiled reference
// (oracle/_ref/libsnn_ref.so). Conv/Depthwise/Pool/BN/... are pinned against hand-derived
// known-answer vectors built from the reference's own unit-test constructions (all-ones inputs,
// pooling sentinels, BN gamma=1 mu=0 var=1 -> 1/sqrt(1.001)) in tests/golden/. The reference's own
// numeric ground truth for those ops is ncnn 20211208, which is not vendored and not available
// offline: "parity vs ncnn unpinned".
//
// Accumulation: fp32, in the shader's loop order (bias first, then ky, kx, ic) — see
// shadertemplate_vk_conv2d.comp:158-274. ORC_ACC_DOUBLE=1 switches to double accumulation for
// tolerance studies.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

// ---------------------------------------------------------------------------------------------
// Activation ids follow conv2dVulkan.cpp:57-71 (0 none, 1 relu, 2 relu6, 3 tanh, 4 sigmoid,
// 5 leakyRelu, 6 SiLU). SiLU is the mathematically correct x*sigmoid(x); the reference's 4-pixel
// conv kernel mis-applies pixel 1's sigmoid to pixels 2-4 (vk_conv2d.comp:336-339) and its CPU
// SiLU is a no-op (cpulayer.h:245-252) — both rejected as bugs.
// ---------------------------------------------------------------------------------------------
static inline float orc_act(float v, int act, float alpha) {
    switch (act) {
    case 1: return std::max(v, 0.0f);
    case 2: return std::min(std::max(v, 0.0f), 6.0f);
    case 3: return std::tanh(v);
    case 4: return 1.0f / (1.0f + std::exp(-v));
    case 5: return std::max(v, v * alpha);
    case 6: return v * 1.0f / (1.0f + std::exp(-v));
    default: return v;
    }
}

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void) n;
#endif
}

// BN as every conv/depthwise/batchnorm shader writes it (vk_conv2d.comp:277-288):
//   s = sqrt(var + 0.001); s = max(s, 0.0001); y = gamma/s * (x - mean) + beta      (eps hard-coded)
static inline float orc_bn(float v, const float* bn, int C, int c) {
    const float gamma = bn[c], beta = bn[C + c], mean = bn[2 * C + c], var = bn[3 * C + c];
    float s = std::sqrt(var + 0.001f);
    s       = std::max(s, 0.0001f);
    return ((gamma / s) * (v - mean)) + beta;
}

// ---------------------------------------------------------------------------------------------
// Output-dimension rules (float math then truncation, genericlayer.cpp:64-90). NB the reference accumulates the
// translation with std::max starting from 0 (genericlayer.cpp:66,73,75), so a NEGATIVE translation is clamped to 0:
// a "valid" 3x3 stride-1 conv keeps its input size (the extra columns read out of range = 0), and a global average
// pool only collapses to 1x1 because the avg-pool reader defaults stride to the pool size (modelparser.cpp:385-391).
// ---------------------------------------------------------------------------------------------
// Conv2D: conv2d.cpp:102-113. offsets = {T,B,L,R}; the translation uses offset[0]+offset[1] for BOTH axes.
int orc_conv_out_dim(int in, int k, int stride, int padT, int padB) {
    float scale = 1 / static_cast<float>(stride);
    float translation;
    if (k % 2 != 0) {
        translation = 1 + (static_cast<float>(padT + padB) - static_cast<float>(k)) / static_cast<float>(stride);
    } else {
        translation = 1 + (static_cast<float>(padT + padB - 1) - static_cast<float>(k)) / static_cast<float>(stride);
    }
    float v = scale * in + std::max(0.0f, translation);
    return (int) (uint32_t) v;
}
// Depthwise: separableconvolution.cpp:77-86 — integer math; width uses T+L, height uses B+R.
int orc_depthwise_out_dim(int in, int k, int stride, int padA, int padB) { return (in - k + padA + padB) / stride + 1; }
// Pools: maxpool2d.cpp:26-35 / avgpool2d.cpp:21-30. valid_like = padding in {"0","valid","none"}.
int orc_pool_out_dim(int in, int k, int stride, int valid_like) {
    float scale = 1.0f / stride;
    float translation;
    if (valid_like)
        translation = 1.0f - (static_cast<float>(k) / static_cast<float>(stride));
    else
        translation = 1.0f - 1.0f / static_cast<float>(stride);
    float v = scale * in + std::max(0.0f, translation);
    return (int) (uint32_t) v;
}
// "same"/"valid" -> offsets {T,B,L,R}: conv2d.cpp:39-74 (even k: top/left = k/2-1).
void orc_same_padding(int k, int is_same, int* offs) {
    offs[0] = offs[1] = offs[2] = offs[3] = 0;
    if (is_same && k > 1) {
        int p   = std::max(k / 2, 1);
        offs[0] = offs[1] = offs[2] = offs[3] = p;
        if (k % 2 == 0) {
            offs[0] -= 1;
            offs[2] -= 1;
        }
    }
}

// Source coordinate under a padding mode (vk_conv2d.comp:168-218, vk_pad.comp:53-66):
// 0/1 constant (out of range -> zero), 2 replicate (clamp), 3 reflect (-i ; 2n-2-i).
static inline int orc_src_coord(int s, int n, int mode) {
    if (mode == 2) return std::min(std::max(s, 0), n - 1);
    if (mode == 3) { // one reflection; a coordinate still outside reads zero like an out-of-range texelFetch
        s = (s < 0) ? -s : s;
        s = (s >= n) ? 2 * n - 2 - s : s;
    }
    return (s >= 0 && s < n) ? s : -1;
}

// ---------------------------------------------------------------------------------------------
// Conv2D k x k (shadertemplate_vk_conv2d.comp:148-347; 1x1 twin vk_conv2d_1x1.comp:68-211;
// host conv2dVulkan.cpp:154-217). x NHWC [N,H,W,IC]; w OIHW flat (modelparser.cpp:641-657);
// bias[OC] or null; bn = {gamma[OC],beta[OC],mean[OC],var[OC]} or null; y NHWC [N,OH,OW,OC].
// pad_x / pad_y are the shader's uPadx / uPady (the host passes offsets[0]=T as x and offsets[2]=L
// as y: conv2dVulkan.cpp:183-184 — SURVEY Q5; callers apply that mapping).
// ---------------------------------------------------------------------------------------------
int orc_conv2d(const float* x, int N, int H, int W, int IC, const float* w, const float* bias, const float* bn, int OC, int k, int stride, int pad_x,
               int pad_y, int pad_mode, int act, float alpha, float* y, int OH, int OW) {
    const bool acc_double = getenv("ORC_ACC_DOUBLE") != nullptr;
    // repack OIHW -> [ky][kx][ic][oc] so the oc loop vectorises while each oc keeps the shader's
    // sequential (ky,kx,ic) accumulation order.
    std::vector<float> wp((size_t) k * k * IC * OC);
#pragma omp parallel for collapse(2) schedule(static)
    for (int o = 0; o < OC; ++o)
        for (int i = 0; i < IC; ++i)
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) wp[(((size_t) ky * k + kx) * IC + i) * OC + o] = w[(((size_t) o * IC + i) * k + ky) * k + kx];

#pragma omp parallel
    {
        std::vector<float> acc(OC);
        std::vector<double> accd(acc_double ? OC : 0);
#pragma omp for collapse(3) schedule(static) // (n, oy, ox): small feature maps (7x7) still give every host core work
        for (int n = 0; n < N; ++n) {
            for (int oy = 0; oy < OH; ++oy) {
                for (int ox = 0; ox < OW; ++ox) {
                    for (int o = 0; o < OC; ++o) acc[o] = bias ? bias[o] : 0.0f;
                    if (acc_double)
                        for (int o = 0; o < OC; ++o) accd[o] = acc[o];
                    if (!acc_double) {
                        // fp32 path, blocked over 16 output channels so the accumulators stay in registers; every output
                        // channel still sums bias, then (ky, kx, ic) in the shader's order.
                        constexpr int OB = 16;
                        int sxs[16], sys_[16];
                        const bool small_k = k <= 16;
                        if (small_k)
                            for (int t = 0; t < k; ++t) {
                                sys_[t] = orc_src_coord(oy * stride - pad_y + t, H, pad_mode);
                                sxs[t]  = orc_src_coord(ox * stride - pad_x + t, W, pad_mode);
                            }
                        for (int ob = 0; ob < OC; ob += OB) {
                            const int nb = std::min(OB, OC - ob);
                            float a[OB];
                            for (int j = 0; j < OB; ++j) a[j] = j < nb ? acc[ob + j] : 0.0f;
                            for (int ky = 0; ky < k; ++ky) {
                                const int sy = small_k ? sys_[ky] : orc_src_coord(oy * stride - pad_y + ky, H, pad_mode);
                                if (sy < 0) continue;
                                for (int kx = 0; kx < k; ++kx) {
                                    const int sx = small_k ? sxs[kx] : orc_src_coord(ox * stride - pad_x + kx, W, pad_mode);
                                    if (sx < 0) continue;
                                    const float* xp = x + (((size_t) n * H + sy) * W + sx) * IC;
                                    const float* wk = wp.data() + ((size_t) ky * k + kx) * IC * OC + ob;
                                    if (nb == OB) {
                                        for (int i = 0; i < IC; ++i) {
                                            const float xv  = xp[i];
                                            const float* wr = wk + (size_t) i * OC;
#pragma omp simd
                                            for (int j = 0; j < OB; ++j) a[j] += wr[j] * xv;
                                        }
                                    } else {
                                        for (int i = 0; i < IC; ++i) {
                                            const float xv  = xp[i];
                                            const float* wr = wk + (size_t) i * OC;
                                            for (int j = 0; j < nb; ++j) a[j] += wr[j] * xv;
                                        }
                                    }
                                }
                            }
                            for (int j = 0; j < nb; ++j) acc[ob + j] = a[j];
                        }
                    }
                    for (int ky = 0; acc_double && ky < k; ++ky) {
                        int sy = orc_src_coord(oy * stride - pad_y + ky, H, pad_mode);
                        if (sy < 0) continue;
                        for (int kx = 0; kx < k; ++kx) {
                            int sx = orc_src_coord(ox * stride - pad_x + kx, W, pad_mode);
                            if (sx < 0) continue;
                            const float* xp = x + (((size_t) n * H + sy) * W + sx) * IC;
                            const float* wk = wp.data() + ((size_t) ky * k + kx) * IC * OC;
                            if (!acc_double) {
                                for (int i = 0; i < IC; ++i) {
                                    const float xv  = xp[i];
                                    const float* wr = wk + (size_t) i * OC;
                                    float* a        = acc.data();
#pragma omp simd
                                    for (int o = 0; o < OC; ++o) a[o] += wr[o] * xv;
                                }
                            } else {
                                for (int i = 0; i < IC; ++i) {
                                    const double xv = xp[i];
                                    const float* wr = wk + (size_t) i * OC;
                                    for (int o = 0; o < OC; ++o) accd[o] += (double) wr[o] * xv;
                                }
                            }
                        }
                    }
                    float* yp = y + (((size_t) n * OH + oy) * OW + ox) * OC;
                    for (int o = 0; o < OC; ++o) {
                        float v = acc_double ? (float) accd[o] : acc[o];
                        if (bn) v = orc_bn(v, bn, OC, o);
                        yp[o] = orc_act(v, act, alpha);
                    }
                }
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Depthwise k x k (shadertemplate_vk_depthwise.comp:64-139; separableconvolutionVulkan.cpp).
// Window clipped to the valid range (:77-78) == zero padding. w is [C][k][k] (the parser's CHW
// mats, modelparser.cpp:827-850).
// ---------------------------------------------------------------------------------------------
int orc_depthwise(const float* x, int N, int H, int W, int C, const float* w, const float* bias, const float* bn, int k, int stride, int pad_x, int pad_y,
                  int act, float alpha, float* y, int OH, int OW) {
#pragma omp parallel for collapse(3) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int oy = 0; oy < OH; ++oy) {
            for (int ox = 0; ox < OW; ++ox) {
                const int s0x = ox * stride - pad_x, s0y = oy * stride - pad_y;
                const int sfx = std::max(0, -s0x), sfy = std::max(0, -s0y);
                const int efx = std::min(k, W - s0x), efy = std::min(k, H - s0y);
                float* yp = y + (((size_t) n * OH + oy) * OW + ox) * C;
                for (int c = 0; c < C; ++c) {
                    float v = bias ? bias[c] : 0.0f;
                    for (int fy = sfy; fy < efy; ++fy)
                        for (int fx = sfx; fx < efx; ++fx) v += w[((size_t) c * k + fy) * k + fx] * x[(((size_t) n * H + s0y + fy) * W + s0x + fx) * C + c];
                    if (bn) v = orc_bn(v, bn, C, c);
                    yp[c] = orc_act(v, act, alpha);
                }
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Max / average pooling (shadertemplate_vk_maxpool2d.comp:42-71, vk_avgpool2d.comp:42-69).
// The Vulkan host forces padT = padL = 0 ("Hack it ... not padding on top left in NCNN",
// maxpool2dVulkan.cpp:54-60; avgpool2dVulkan.cpp:55-60): window origin is o*stride, taps clipped
// to the input; max starts at -100000; avg divides by the number of valid taps.
// ---------------------------------------------------------------------------------------------
int orc_pool2d(const float* x, int N, int H, int W, int C, int k, int stride, int is_avg, float* y, int OH, int OW) {
#pragma omp parallel for collapse(3) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int oy = 0; oy < OH; ++oy) {
            for (int ox = 0; ox < OW; ++ox) {
                const int sx = ox * stride, sy = oy * stride;
                const int efx = std::min(k, W - sx), efy = std::min(k, H - sy);
                float* yp = y + (((size_t) n * OH + oy) * OW + ox) * C;
                for (int c = 0; c < C; ++c) {
                    float v = is_avg ? 0.0f : -100000.0f, num = 0.0f;
                    for (int fy = 0; fy < efy; ++fy)
                        for (int fx = 0; fx < efx; ++fx) {
                            float t = x[(((size_t) n * H + sy + fy) * W + sx + fx) * C + c];
                            if (is_avg) {
                                v += t;
                                num += 1.0f;
                            } else {
                                v = std::max(v, t);
                            }
                        }
                    yp[c] = is_avg ? v / num : v;
                }
            }
        }
    }
    return 0;
}

// Add + activation (shadertemplate_vk_add.comp:41-90).
int orc_add(const float* a, const float* b, size_t count, int act, float alpha, float* y) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; ++i) y[i] = orc_act(a[i] + b[i], act, alpha);
    return 0;
}

// Standalone BatchNormalization + activation (shadertemplate_vk_batchnorm.comp:54-69).
int orc_batchnorm(const float* x, size_t pixels, int C, const float* bn, int act, float alpha, float* y) {
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p)
        for (int c = 0; c < C; ++c) y[p * C + c] = orc_act(orc_bn(x[p * C + c], bn, C, c), act, alpha);
    return 0;
}

// Standalone activation (shadertemplate_vk_activation.comp:41-85).
int orc_activation(const float* x, size_t count, int act, float alpha, float* y) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < count; ++i) y[i] = orc_act(x[i], act, alpha);
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Dense (cpulayer.h:136-171): y = W*x + b with W the flat JSON kernel viewed row-major [out][in]
// (cpulayer.h:162), then activation (cpulayer.h:199-261). Activation strings are the CPU
// map's (cpulayer.h:38-42): "relu","leakyRelu","sigmoid","softmax","tanh","SiLU"(no-op there; we
// keep it a no-op HERE because this function restates the CPU path exactly),"identity"/"".
// act ids for this entry: 0 identity, 1 relu, 3 tanh, 4 sigmoid, 5 leakyRelu, 6 SiLU(no-op), 7 softmax.
// x is [N][n_in]; y is [N][n_out].
// ---------------------------------------------------------------------------------------------
static void orc_softmax_row(float* v, int n) {
    // cpulayer.h:175-191: max-subtracted, exp in float, float accumulate
    float mx = -FLT_MAX;
    for (int i = 0; i < n; ++i) mx = std::max(mx, v[i]);
    for (int i = 0; i < n; ++i) v[i] = std::exp(v[i] - mx);
    float div = 0.0f;
    for (int i = 0; i < n; ++i) div += v[i];
    for (int i = 0; i < n; ++i) v[i] = v[i] / div;
}

int orc_dense(const float* x, int N, int n_in, const float* kernel, const float* bias, int n_out, int act, float alpha, float* y) {
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const float* xr = x + (size_t) n * n_in;
        float* yr       = y + (size_t) n * n_out;
        for (int o = 0; o < n_out; ++o) {
            const float* wr = kernel + (size_t) o * n_in;
            float acc       = 0.0f;
            for (int i = 0; i < n_in; ++i) acc += wr[i] * xr[i];
            yr[o] = acc + (bias ? bias[o] : 0.0f);
        }
        switch (act) {
        case 1:
            for (int o = 0; o < n_out; ++o) yr[o] = yr[o] > 0 ? yr[o] : 0.0f * yr[o]; // leakyRelu(val, 0.0), cpulayer.h:205
            break;
        case 5:
            for (int o = 0; o < n_out; ++o) yr[o] = yr[o] > 0 ? yr[o] : alpha * yr[o];
            break;
        case 4:
            for (int o = 0; o < n_out; ++o) yr[o] = 1.0f / (1.0f + std::exp(-yr[o]));
            break;
        case 3:
            for (int o = 0; o < n_out; ++o) yr[o] = (std::exp(2 * yr[o]) - 1) / (std::exp(2 * yr[o]) + 1); // cpulayer.h:195
            break;
        case 7: orc_softmax_row(yr, n_out); break;
        default: break; // identity; SiLU is a by-value no-op in the reference CPU code
        }
    }
    return 0;
}

// Softmax over the last dim of [rows][n] (cpulayer.h:175-191).
int orc_softmax(const float* x, int rows, int n, float* y) {
    for (int r = 0; r < rows; ++r) {
        std::memcpy(y + (size_t) r * n, x + (size_t) r * n, sizeof(float) * n);
        orc_softmax_row(y + (size_t) r * n, n);
    }
    return 0;
}

// Classifier index: argmax + 1 (core.cpp:228-233; 1-based). First maximum wins
// (std::max_element semantics).
int orc_argmax1(const float* x, int rows, int n, int* idx) {
    for (int r = 0; r < rows; ++r) {
        const float* v = x + (size_t) r * n;
        int best       = 0;
        for (int i = 1; i < n; ++i)
            if (v[i] > v[best]) best = i;
        idx[r] = best + 1;
    }
    return 0;
}

// Flatten, CPU flavour: HWC order (cpulayer.h:94-115) — for NHWC input this is a copy.
int orc_flatten(const float* x, size_t count, float* y) {
    std::memcpy(y, x, count * sizeof(float));
    return 0;
}

// Concatenate along channels (concatenation.h:25-40, vk_concat.comp:39-52).
int orc_concat(const float* a, int Ca, const float* b, int Cb, size_t pixels, float* y) {
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < pixels; ++p) {
        std::memcpy(y + p * (Ca + Cb), a + p * Ca, sizeof(float) * Ca);
        std::memcpy(y + p * (Ca + Cb) + Ca, b + p * Cb, sizeof(float) * Cb);
    }
    return 0;
}

// UpSampling2D (upsampling2d.h:26-47). nearest: src = clamp(floor(dst * (1/scale)))
// (vk_upsampling2d_nearest.comp:43-64); bilinear: half-pixel centres, clamp to [0, n-1], taps
// x11/x12=x11+1 (vk_upsampling2d_bilinear.comp:43-86; an out-of-range x12 has weight 0).
int orc_upsample(const float* x, int N, int H, int W, int C, float scale, int bilinear, float* y, int OH, int OW) {
    const float inv = 1.0f / scale;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int oy = 0; oy < OH; ++oy) {
            for (int ox = 0; ox < OW; ++ox) {
                float* yp = y + (((size_t) n * OH + oy) * OW + ox) * C;
                if (!bilinear) {
                    int x1          = std::min(std::max((int) std::floor((float) ox * inv), 0), W - 1);
                    int y1          = std::min(std::max((int) std::floor((float) oy * inv), 0), H - 1);
                    const float* xp = x + (((size_t) n * H + y1) * W + x1) * C;
                    for (int c = 0; c < C; ++c) yp[c] = xp[c];
                } else {
                    float offs = 0.5f - 0.5f * inv;
                    float sx   = std::min(std::max((float) ox * inv - offs, 0.0f), (float) (W - 1));
                    float sy   = std::min(std::max((float) oy * inv - offs, 0.0f), (float) (H - 1));
                    int x11 = (int) std::floor(sx), x12 = x11 + 1;
                    int y11 = (int) std::floor(sy), y12 = y11 + 1;
                    auto fetch = [&](int xx, int yy, int c) -> float {
                        if (xx < 0 || xx >= W || yy < 0 || yy >= H) return 0.0f;
                        return x[(((size_t) n * H + yy) * W + xx) * C + c];
                    };
                    for (int c = 0; c < C; ++c) {
                        float r1 = fetch(x11, y11, c), r2 = fetch(x12, y11, c), r3 = fetch(x12, y12, c), r4 = fetch(x11, y12, c);
                        yp[c] = r1 * (((float) x12 - sx) * ((float) y12 - sy)) + r2 * ((sx - (float) x11) * ((float) y12 - sy)) +
                                r3 * ((sx - (float) x11) * (sy - (float) y11)) + r4 * (((float) x12 - sx) * (sy - (float) y11));
                    }
                }
            }
        }
    }
    return 0;
}

// Pad (padlayer.cpp:26-70 dims; vk_pad.comp:42-70): mode 0/1 constant(0), 2 replicate, 3 reflect.
// The shader's uPad = (x: left? ...) — host passes {T? L?}: padlayerVulkan passes offsets so that
// pos.xy - uPad; we take explicit pad_x (left) and pad_y (top).
int orc_pad(const float* x, int N, int H, int W, int C, int pad_x, int pad_y, int mode, float* y, int OH, int OW) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int oy = 0; oy < OH; ++oy) {
            for (int ox = 0; ox < OW; ++ox) {
                int sx    = orc_src_coord(ox - pad_x, W, mode);
                int sy    = orc_src_coord(oy - pad_y, H, mode);
                float* yp = y + (((size_t) n * OH + oy) * OW + ox) * C;
                if (sx < 0 || sy < 0) {
                    for (int c = 0; c < C; ++c) yp[c] = 0.0f;
                } else {
                    const float* xp = x + (((size_t) n * H + sy) * W + sx) * C;
                    for (int c = 0; c < C; ++c) yp[c] = xp[c];
                }
            }
        }
    }
    return 0;
}

// InstanceNorm (instancenorm.h:28-54, vk_instancenorm.comp:53-175): per (n,c) mean and BIASED
// variance over H*W, eps hard-coded 1e-5 (:128), y = (x-mean)*gamma/sqrt(var+eps)+beta, act.
// The shader's reduction order (strided partial sums over a 2-D workgroup) is hardware-shaped; the
// oracle accumulates in double and rounds once — the parity tolerance covers the difference.
int orc_instancenorm(const float* x, int N, int H, int W, int C, const float* gamma, const float* beta, int act, float alpha, float* y) {
    const size_t HW = (size_t) H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int c = 0; c < C; ++c) {
            const float* xp = x + (size_t) n * HW * C + c;
            double s        = 0.0;
            for (size_t p = 0; p < HW; ++p) s += xp[p * C];
            const float mean = (float) (s / (double) HW);
            double v         = 0.0;
            for (size_t p = 0; p < HW; ++p) {
                double d = (double) xp[p * C] - (double) mean;
                v += d * d;
            }
            const float var   = (float) (v / (double) HW);
            const float sigma = std::sqrt(var + 0.00001f);
            const float mul   = gamma[c] / sigma;
            float* yp         = y + (size_t) n * HW * C + c;
            for (size_t p = 0; p < HW; ++p) yp[p * C] = orc_act((xp[p * C] - mean) * mul + beta[c], act, alpha);
        }
    }
    return 0;
}

// Subpixel / depth_to_space(r) + tanh ALWAYS (subpixelmerge.h:26-47; component = x%r + r*(y%r),
// fs_subpixel.glsl:41; vk_subpixel.comp:58-66 applies tanh unconditionally — SURVEY Q9; Keras
// semantics per demo/modelInferenceESPCN.py:36-38,65-67). Input [N,H,W,r*r] -> output [N,H*r,W*r,1].
int orc_subpixel(const float* x, int N, int H, int W, int r, float* y) {
    const int OH = H * r, OW = W * r, C = r * r;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n)
        for (int oy = 0; oy < OH; ++oy)
            for (int ox = 0; ox < OW; ++ox) {
                int comp                              = (ox % r) + r * (oy % r);
                float v                               = x[(((size_t) n * H + oy / r) * W + ox / r) * C + comp];
                y[((size_t) n * OH + oy) * OW + ox] = std::tanh(v);
            }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// YOLOv3-tiny decode + NMS (yololayer.cpp:31-38 constants, :102-164 decode, :56-100 IoU/NMS,
// :177-226 driver). heads[i] is one image's head i in NHWC with channel pitch Cp (the reference
// reads a C4-padded texture: 18 -> 20, :160). Output rows {class, score, x, y, w, h}; returns count.
// Quirk kept: score = 1/(1 + e^-obj * (1 + e^-cls)) exactly as parenthesised at :136.
// ---------------------------------------------------------------------------------------------
struct OrcBox {
    int cls;
    float score, x, y, w, h;
};
static float orc_iou(const OrcBox& a, const OrcBox& b) {
    float ix0 = std::max(a.x, b.x), iy0 = std::max(a.y, b.y);
    float ix1 = std::min(a.x + a.w, b.x + b.w), iy1 = std::min(a.y + a.h, b.y + b.h);
    if (ix1 < ix0 || iy1 < iy0) return 0;
    float a0 = a.w * a.h, a1 = b.w * b.h, ai = (ix1 - ix0) * (iy1 - iy0);
    return ai / (a0 + a1 - ai);
}
int orc_yolo(const float* head0, const float* head1, int Cp, int net_w, int net_h, float conf_thresh, float iou_thresh, float* out, int max_out) {
    static const int gridScale[2] = {32, 16};
    static const float anchors[]  = {10, 14, 23, 27, 37, 58, 81, 82, 135, 169, 344, 319};
    static const float masks[]    = {3, 4, 5, 1, 2, 3};
    const int GC = 3, NCLS = 1, NFIX = 5, ONUM = NCLS + NFIX;
    std::vector<OrcBox> boxes;
    const float* heads[2] = {head0, head1};
    for (int yi = 0; yi < 2; ++yi) {
        const int gw = net_w / gridScale[yi], gh = net_h / gridScale[yi];
        const float* data = heads[yi];
        const int netW = (int) ((float) gridScale[yi] * gw), netH = (int) ((float) gridScale[yi] * gh);
        for (int gy = 0; gy < gh; ++gy)
            for (int gx = 0; gx < gw; ++gx) {
                const float* px = data + ((size_t) gy * gw + gx) * Cp;
                for (int gc = 0; gc < GC; ++gc) {
                    const float* d = px + gc * ONUM;
                    int cls        = 0;
                    float maxLogit = -FLT_MAX;
                    for (int i = NFIX; i < ONUM; ++i)
                        if (d[i] > maxLogit) {
                            maxLogit = d[i];
                            cls      = i - NFIX;
                        }
                    int ai     = (int) masks[gc + yi * GC];
                    float bw = anchors[ai * 2], bh = anchors[ai * 2 + 1];
                    float prob = 1.f / ((1.f + std::exp(-d[4]) * (1.f + std::exp(-maxLogit))));
                    if (prob > conf_thresh) {
                        float cx = (gx + 1.0f / (1.0f + std::exp(-d[0]))) / gw;
                        float cy = (gy + 1.0f / (1.0f + std::exp(-d[1]))) / gh;
                        float w_ = std::exp(d[2]) * bw / netW;
                        float h_ = std::exp(d[3]) * bh / netH;
                        boxes.push_back({cls, prob, cx - w_ / 2, cy - h_ / 2, w_, h_});
                    }
                }
            }
    }
    std::stable_sort(boxes.begin(), boxes.end(), [](const OrcBox& l, const OrcBox& r) { return l.score > r.score; });
    std::vector<char> merged(boxes.size(), 0);
    int cnt = 0;
    for (size_t i = 0; i < boxes.size(); ++i) {
        if (merged[i]) continue;
        for (size_t j = i + 1; j < boxes.size(); ++j) {
            if (merged[j] || boxes[i].cls != boxes[j].cls) continue;
            if (orc_iou(boxes[i], boxes[j]) > iou_thresh) merged[j] = 1;
        }
        if (cnt < max_out) {
            float* o = out + (size_t) cnt * 6;
            o[0] = (float) boxes[i].cls, o[1] = boxes[i].score, o[2] = boxes[i].x, o[3] = boxes[i].y, o[4] = boxes[i].w, o[5] = boxes[i].h;
        }
        ++cnt;
    }
    return std::min(cnt, max_out);
}

// ---------------------------------------------------------------------------------------------
// Layout converters for the API edge (shaderUnitTest.cpp:87-131 hwcToC4):
// element (x,y,c) of a C4HW4 texture lives at (((c/4)*H + y)*W + x)*4 + c%4.
// ---------------------------------------------------------------------------------------------
int orc_hwc_to_c4hw4(const float* hwc, int H, int W, int C, float* c4) {
    const int D = (C + 3) / 4;
    std::memset(c4, 0, sizeof(float) * (size_t) D * H * W * 4);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) c4[((((size_t) (c / 4)) * H + y) * W + x) * 4 + (c % 4)] = hwc[((size_t) y * W + x) * C + c];
    return 0;
}
int orc_c4hw4_to_hwc(const float* c4, int H, int W, int C, float* hwc) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) hwc[((size_t) y * W + x) * C + c] = c4[((((size_t) (c / 4)) * H + y) * W + x) * 4 + (c % 4)];
    return 0;
}

// ---------------------------------------------------------------------------------------------
// The unit-test PRNG (demo/common/prng.h:52-88, Bob Adolf's public-domain additive lagged
// Fibonacci generator, lags 24/55 over a 64-entry ring, refilled every 55 draws by running
// 550 steps forward; seeded "i*2147483647+seed" then run 10000 draws). Every reference op test
// seeds it with 7767517 (convolutionTest.cpp:417) and draws weights with RandomFloat(-1.2,1.2)
// (testutil.cpp:40-47). Restated so golden vectors can be regenerated without the reference tree.
// ---------------------------------------------------------------------------------------------
struct OrcPrng {
    uint64_t s[64];
    unsigned i, c;
};
static OrcPrng g_prng;
static uint64_t orc_prng_next(OrcPrng* st) {
    unsigned steps;
    if (!st->c) {
        steps = (55 * 10 - 55) + 1;
        st->c = 55 - 1;
    } else {
        steps = 1;
        st->c--;
    }
    unsigned i = 0;
    for (unsigned r = 0; r < steps; ++r) {
        i             = st->i;
        st->s[i & 63] = st->s[(i + 64 - 24) & 63] + st->s[(i + 64 - 55) & 63];
        st->i         = (st->i + 1) & 0xFFFF; // the reference's counter is a uint_fast16_t; only i&63 is ever used
    }
    return st->s[i & 63];
}
void orc_srand(uint64_t seed) {
    g_prng.c    = 55;
    g_prng.i    = 0;
    g_prng.s[0] = seed;
    for (unsigned i = 1; i < 64; ++i) g_prng.s[i] = i * UINT64_C(2147483647) + seed;
    for (int i = 0; i < 10000; ++i) orc_prng_next(&g_prng);
}
uint64_t orc_rand_u64() { return orc_prng_next(&g_prng); }
float orc_random_float(float a, float b) {
    float random = ((float) orc_prng_next(&g_prng)) / (float) uint64_t(-1);
    float diff   = b - a;
    volatile float rd = random * diff; // two roundings as in the reference build (no FMA contraction): bit-exact stream
    return a + rd;
}
void orc_random_fill(float* dst, size_t n, float a, float b) {
    for (size_t i = 0; i < n; ++i) dst[i] = orc_random_float(a, b);
}

// The reference's comparator (demo/common/testutil.cpp:351-361): equal, or |a-b| <= eps, or
// |a-b| < eps*max(|a|,|b|). Returns the number of mismatching elements.
size_t orc_compare(const float* a, const float* b, size_t n, float eps) {
    size_t bad = 0;
    for (size_t i = 0; i < n; ++i) {
        if (a[i] == b[i]) continue;
        float diff = std::fabs(a[i] - b[i]);
        if (diff <= eps) continue;
        if (diff < eps * std::max(std::fabs(a[i]), std::fabs(b[i]))) continue;
        if (std::isnan(a[i]) && std::isnan(b[i])) continue;
        ++bad;
    }
    return bad;
}

} // extern "C"
