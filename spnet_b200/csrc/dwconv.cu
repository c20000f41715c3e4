// Depthwise 3x3, stride 1, 'same' padding, NHWC — forward, data-gradient and
// weight-gradient. This is the depthwise half of keras SeparableConv2D as used by
// keras.applications.Xception (reference call site spnet/models.py:359).
//
// HBM-bound: every thread owns one 16-byte channel vector (8 bf16 / 4 fp32) of one
// pixel column and slides down a strip of rows, so each input vector is loaded
// from global memory once per (column, +-1 neighbour) and all arithmetic is fp32
// in registers. Xception is pre-activation (ReLU -> sepconv -> BN), and in
// training the producer's BatchNorm can only be applied once its batch statistics
// exist, so BN-apply (per-channel affine) + ReLU are fused on the LOAD side here;
// zero padding is applied after that transform, as TF does.
#include "common.cuh"

namespace {

struct DwEpilogue {
    // All optional (nullptr = off). Used by the data-gradient pass.
    const void* mask_src;    // [B,H,W,C]  multiply result by (mask_a*src+mask_b > 0)
    const float* mask_a;     // [C] or nullptr (=1)
    const float* mask_b;     // [C] or nullptr (=0)
    const void* add_src;     // [B,H,W,C]  added after masking
    const void* add_strided; // [B,ceil(H/2),ceil(W/2),C] added at even (h,w) after masking
};

template <typename T, bool AFFINE, bool RELU, bool FLIP>
__global__ void __launch_bounds__(256) dw3x3_kernel(const T* __restrict__ in, const float* __restrict__ k,
                                                    const float* __restrict__ in_a,
                                                    const float* __restrict__ in_b, T* __restrict__ out, int B,
                                                    int H, int W, int C, int R, int nstrips, DwEpilogue ep) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const long long total = (long long)B * nstrips * W * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int cv = (int)(idx % CV);
    long long r = idx / CV;
    const int w = (int)(r % W);
    r /= W;
    const int strip = (int)(r % nstrips);
    const int b = (int)(r / nstrips);
    const int c0 = cv * V;
    const int h0 = strip * R;
    const int h1 = min(H, h0 + R);

    float wt[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int ts = FLIP ? 8 - t : t;
#pragma unroll
        for (int i = 0; i < V; ++i) wt[t][i] = k[ts * C + c0 + i];
    }
    float a[V], bb[V];
    if (AFFINE) {
#pragma unroll
        for (int i = 0; i < V; ++i) { a[i] = in_a[c0 + i]; bb[i] = in_b[c0 + i]; }
    }
    float accA[V], accB[V], accC[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { accA[i] = 0.f; accB[i] = 0.f; accC[i] = 0.f; }

    const size_t img = (size_t)b * H * W;
    for (int ih = h0 - 1; ih <= h1; ++ih) {
        const bool rowok = ih >= 0 && ih < H;
        float x[3][V];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
            const int iw = w - 1 + kw;
            if (rowok && iw >= 0 && iw < W) {
                load_vec(in + (img + (size_t)ih * W + iw) * C + c0, x[kw]);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float v = x[kw][i];
                    if (AFFINE) v = fmaf(v, a[i], bb[i]);
                    if (RELU) v = fmaxf(v, 0.f);
                    x[kw][i] = v;
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) x[kw][i] = 0.f;
            }
        }
        // input row ih is tap row 2 of output ih-1, row 1 of output ih, row 0 of output ih+1
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                accA[i] = fmaf(x[kw][i], wt[6 + kw][i], accA[i]);
                accB[i] = fmaf(x[kw][i], wt[3 + kw][i], accB[i]);
                accC[i] = fmaf(x[kw][i], wt[0 + kw][i], accC[i]);
            }
        }
        const int oh = ih - 1;
        if (oh >= h0 && oh < h1) {
            const size_t o = (img + (size_t)oh * W + w) * C + c0;
            if (ep.mask_src) {
                float m[V];
                load_vec(reinterpret_cast<const T*>(ep.mask_src) + o, m);
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float v = m[i];
                    if (ep.mask_a) v = fmaf(v, ep.mask_a[c0 + i], ep.mask_b[c0 + i]);
                    if (!(v > 0.f)) accA[i] = 0.f;
                }
            }
            if (ep.add_src) {
                float m[V];
                load_vec(reinterpret_cast<const T*>(ep.add_src) + o, m);
#pragma unroll
                for (int i = 0; i < V; ++i) accA[i] += m[i];
            }
            if (ep.add_strided && ((oh | w) & 1) == 0) {
                const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
                float m[V];
                load_vec(reinterpret_cast<const T*>(ep.add_strided) +
                             (((size_t)b * H2 + (oh >> 1)) * W2 + (w >> 1)) * C + c0, m);
#pragma unroll
                for (int i = 0; i < V; ++i) accA[i] += m[i];
            }
            store_vec(out + o, accA);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) { accA[i] = accB[i]; accB[i] = accC[i]; accC[i] = 0.f; }
    }
}

// Weight gradient: dk[kh,kw,c] = sum_{b,h,w} act(in)[b,h+kh-1,w+kw-1,c] * g[b,h,w,c].
// Persistent CTAs: blockDim = cvb*kcols with a fixed channel vector per thread, so the 9*V
// partial sums stay in registers across the whole grid-stride loop; one shared-memory
// reduction and 9*C global atomics per CTA at the end.
template <typename T, bool AFFINE, bool RELU>
__global__ void __launch_bounds__(256) dw3x3_wgrad_kernel(const T* __restrict__ in, const T* __restrict__ g,
                                                          const float* __restrict__ in_a,
                                                          const float* __restrict__ in_b,
                                                          float* __restrict__ dk, int B, int H, int W, int C,
                                                          int R, int nstrips, int cvb, int kcols) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int cvl = threadIdx.x % cvb;
    const int col = threadIdx.x / cvb;
    const int cv = blockIdx.y * cvb + cvl;
    const bool active = cv < CV;
    const int c0 = cv * V;

    float acc[9][V];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[t][i] = 0.f;

    if (active) {
        float a[V], bb[V];
        if (AFFINE) {
#pragma unroll
            for (int i = 0; i < V; ++i) { a[i] = in_a[c0 + i]; bb[i] = in_b[c0 + i]; }
        }
        const long long items = (long long)B * nstrips * W;
        for (long long it = (long long)blockIdx.x * kcols + col; it < items; it += (long long)gridDim.x * kcols) {
            const int w = (int)(it % W);
            long long r = it / W;
            const int strip = (int)(r % nstrips);
            const int b = (int)(r / nstrips);
            const int h0 = strip * R;
            const int h1 = min(H, h0 + R);
            const size_t img = (size_t)b * H * W;
            // g rows ih-1, ih, ih+1 (zero outside the strip: other strips own those outputs)
            float gm[V], g0[V], gp[V];
#pragma unroll
            for (int i = 0; i < V; ++i) { gm[i] = 0.f; g0[i] = 0.f; gp[i] = 0.f; }
            if (h0 < h1) load_vec(g + (img + (size_t)h0 * W + w) * C + c0, gp);  // row h0 = (h0-1)+1
            for (int ih = h0 - 1; ih <= h1; ++ih) {
                const bool rowok = ih >= 0 && ih < H;
                if (rowok) {
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) {
                        const int iw = w - 1 + kw;
                        if (iw < 0 || iw >= W) continue;
                        float x[V];
                        load_vec(in + (img + (size_t)ih * W + iw) * C + c0, x);
#pragma unroll
                        for (int i = 0; i < V; ++i) {
                            float v = x[i];
                            if (AFFINE) v = fmaf(v, a[i], bb[i]);
                            if (RELU) v = fmaxf(v, 0.f);
                            // in row ih pairs with g row ih+1 (kh=0), ih (kh=1), ih-1 (kh=2)
                            acc[0 + kw][i] = fmaf(v, gp[i], acc[0 + kw][i]);
                            acc[3 + kw][i] = fmaf(v, g0[i], acc[3 + kw][i]);
                            acc[6 + kw][i] = fmaf(v, gm[i], acc[6 + kw][i]);
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < V; ++i) { gm[i] = g0[i]; g0[i] = gp[i]; }
                const int nh = ih + 2;
                if (nh >= h0 && nh < h1) {
                    load_vec(g + (img + (size_t)nh * W + w) * C + c0, gp);
                } else {
#pragma unroll
                    for (int i = 0; i < V; ++i) gp[i] = 0.f;
                }
            }
        }
    }
    // reduce over the kcols threads that share a channel vector, tap by tap
    __shared__ float red[256 * 8];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < V; ++i) red[threadIdx.x * V + i] = acc[t][i];
        __syncthreads();
        if (col == 0 && active) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float s = 0.f;
                for (int j = 0; j < kcols; ++j) s += red[(j * cvb + cvl) * V + i];
                atomicAdd(dk + (size_t)t * C + c0 + i, s);
            }
        }
    }
}

static void pick_strips(int B, int H, int W, int CV, int* R, int* nstrips) {
    int r = H < 32 ? H : 32;
    // keep at least ~4 waves of 256-thread CTAs on 148 SMs when the tensor allows it
    while (r > 4 && (long long)B * ceil_div(H, r) * W * CV < 600000) r = (r + 1) / 2;
    *R = r;
    *nstrips = ceil_div(H, r);
}

template <typename T>
int launch_dw(const void* in, const float* k, const float* in_a, const float* in_b, int relu, int flip,
              void* out, int B, int H, int W, int C, DwEpilogue ep, cudaStream_t stream) {
    constexpr int V = VecN<T>::N;
    int R, nstrips;
    pick_strips(B, H, W, C / V, &R, &nstrips);
    const long long total = (long long)B * nstrips * W * (C / V);
    const int grid = ceil_div(total, 256);
    const T* x = reinterpret_cast<const T*>(in);
    T* y = reinterpret_cast<T*>(out);
#define DW_LAUNCH(AF, RL, FL) \
    dw3x3_kernel<T, AF, RL, FL><<<grid, 256, 0, stream>>>(x, k, in_a, in_b, y, B, H, W, C, R, nstrips, ep)
    if (flip) { DW_LAUNCH(false, false, true); }
    else if (in_a && relu) { DW_LAUNCH(true, true, false); }
    else if (in_a) { DW_LAUNCH(true, false, false); }
    else if (relu) { DW_LAUNCH(false, true, false); }
    else { DW_LAUNCH(false, false, false); }
#undef DW_LAUNCH
    return spnet_check_launch("dw3x3");
}

template <typename T>
int launch_dw_wgrad(const void* in, const void* g, const float* in_a, const float* in_b, int relu, float* dk,
                    int B, int H, int W, int C, cudaStream_t stream) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    int R, nstrips;
    pick_strips(B, H, W, CV, &R, &nstrips);
    const int nchunks = ceil_div(CV, 128);
    const int cvb = ceil_div(CV, nchunks);
    int kcols = 256 / cvb;
    if (kcols < 1) kcols = 1;
    const long long items = (long long)B * nstrips * W;
    int gx = ceil_div(items, kcols);
    const int cap = (2 * 148 + nchunks - 1) / nchunks;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, nchunks);
    const T* x = reinterpret_cast<const T*>(in);
    const T* gg = reinterpret_cast<const T*>(g);
#define DWW_LAUNCH(AF, RL) \
    dw3x3_wgrad_kernel<T, AF, RL><<<grid, cvb * kcols, 0, stream>>>(x, gg, in_a, in_b, dk, B, H, W, C, R, nstrips, cvb, kcols)
    if (in_a && relu) { DWW_LAUNCH(true, true); }
    else if (in_a) { DWW_LAUNCH(true, false); }
    else if (relu) { DWW_LAUNCH(false, true); }
    else { DWW_LAUNCH(false, false); }
#undef DWW_LAUNCH
    return spnet_check_launch("dw3x3_wgrad");
}

int check_dw_args(const char* who, const void* in, const void* out, int dtype, int B, int H, int W, int C) {
    SPNET_REQUIRE(in && out, "%s: null pointer", who);
    SPNET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", who);
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_REQUIRE(C % V == 0, "%s: C=%d must be a multiple of %d", who, C, V);
    return SPNET_OK;
}

}  // namespace

extern "C" {

// out = dw3x3(act(in)),  act(v) = relu?(in_a*v + in_b)   (in_a/in_b nullable, fp32 [C])
// k: [3,3,C] fp32 (keras depthwise_kernel (3,3,C,1) flattened)
int spnet_dwconv3x3_fwd(const void* in, const float* k, const float* in_a, const float* in_b, int relu,
                        void* out, int dtype, int B, int H, int W, int C, cudaStream_t stream) {
    int rc = check_dw_args("dwconv3x3_fwd", in, out, dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(k && ((in_a == nullptr) == (in_b == nullptr)), "dwconv3x3_fwd: bad weight/affine pointers");
    DwEpilogue ep = {nullptr, nullptr, nullptr, nullptr, nullptr};
    SPNET_DISPATCH_DTYPE(dtype, return launch_dw<T>(in, k, in_a, in_b, relu, 0, out, B, H, W, C, ep, stream));
}

// gin = dw3x3^T(gout) [* (mask_a*mask_src+mask_b > 0)] [+ add_src] [+ add_strided at even (h,w)]
int spnet_dwconv3x3_dgrad(const void* gout, const float* k, void* gin, const void* mask_src,
                          const float* mask_a, const float* mask_b, const void* add_src,
                          const void* add_strided, int dtype, int B, int H, int W, int C,
                          cudaStream_t stream) {
    int rc = check_dw_args("dwconv3x3_dgrad", gout, gin, dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(k && ((mask_a == nullptr) == (mask_b == nullptr)), "dwconv3x3_dgrad: bad pointers");
    SPNET_REQUIRE(mask_src || !mask_a, "dwconv3x3_dgrad: mask affine without mask_src");
    DwEpilogue ep = {mask_src, mask_a, mask_b, add_src, add_strided};
    SPNET_DISPATCH_DTYPE(dtype, return launch_dw<T>(gout, k, nullptr, nullptr, 0, 1, gin, B, H, W, C, ep, stream));
}

// dk[3,3,C] += sum act(in) (*) gout     (dk fp32, accumulated with atomics: zero it first)
int spnet_dwconv3x3_wgrad(const void* in, const void* gout, const float* in_a, const float* in_b, int relu,
                          float* dk, int dtype, int B, int H, int W, int C, cudaStream_t stream) {
    int rc = check_dw_args("dwconv3x3_wgrad", in, gout, dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(dk && ((in_a == nullptr) == (in_b == nullptr)), "dwconv3x3_wgrad: bad pointers");
    SPNET_DISPATCH_DTYPE(dtype, return launch_dw_wgrad<T>(in, gout, in_a, in_b, relu, dk, B, H, W, C, stream));
}

}  // extern "C"
