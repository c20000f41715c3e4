// MaxPooling2D(3, strides 2, 'same') fused with the BatchNorm-apply of both branches and
// the residual add of the Xception entry/exit blocks (keras.applications.Xception blocks
// 2-4 and 13; reference call site spnet/models.py:359), its backward, and the stride-2
// row/column subsample that feeds the 1x1 stride-2 residual convolutions.
//
// TF 'SAME' padding is asymmetric: out = ceil(in/2), pad_total = max((out-1)*2+3-in, 0),
// pad_before = pad_total/2 (even sizes pad only at the end). Padded cells never win the max.
#include "common.cuh"

namespace {

__host__ __device__ inline int same_pad_before(int in) {
    const int out = (in + 1) / 2;
    int total = (out - 1) * 2 + 3 - in;
    if (total < 0) total = 0;
    return total / 2;
}

// per-channel coefficients of one channel vector with 16-byte loads
template <int V> __device__ __forceinline__ void load_coef(const float* __restrict__ p, int c0, float (&v)[V]) {
#pragma unroll
    for (int i = 0; i < V; i += 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(p + c0 + i));
        v[i] = q.x; v[i + 1] = q.y; v[i + 2] = q.z; v[i + 3] = q.w;
    }
}

// Both pooling kernels: blockIdx.x = one row of the tensor the kernel WRITES (image, output row), the threads
// walk that row's (column, channel vector) items; the image / row split is per CTA and one division per item
// is left (the one-thread-per-item versions spent three 32-bit divisions and eight single-byte argmax stores
// per item and ran at 30-40 % of the HBM roofline).
//
// out = max_{3x3,s2}(a*z+b) + (ra*res+rb);  argmax (optional) = kh*3+kw of the first maximum
template <typename T>
__global__ void __launch_bounds__(256) maxpool_add_fwd_kernel(const T* __restrict__ z, const float* __restrict__ a,
                                                              const float* __restrict__ b,
                                                              const T* __restrict__ res,
                                                              const float* __restrict__ ra,
                                                              const float* __restrict__ rb, T* __restrict__ out,
                                                              unsigned char* __restrict__ argmax, int B, int H,
                                                              int W, int C, int OH, int OW, int pt, int pl) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int row = blockIdx.x;  // bi * OH + oh
    const int bi = row / OH, oh = row - bi * OH;
    const int items = OW * CV;
    // the three input rows of this output row (clamped; flagged when they are padding)
    int ihc[3];
    bool okh[3];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
        const int ih = oh * 2 - pt + kh;
        ihc[kh] = min(max(ih, 0), H - 1);
        okh[kh] = ih == ihc[kh];
    }
    for (int it = blockIdx.y * blockDim.x + threadIdx.x; it < items; it += blockDim.x * gridDim.y) {
        const int ow = it / CV, cv = it - ow * CV;
        const int c0 = cv * V;
        const size_t idx = (size_t)row * items + it;
        float av[V], bv[V], best[V];
        int arg[V];
        if (a) {
            load_coef<V>(a, c0, av);
            load_coef<V>(b, c0, bv);
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
            if (!a) { av[i] = 1.f; bv[i] = 0.f; }
            best[i] = -INFINITY;
            arg[i] = 0;
        }
        // all nine loads are issued before the first compare (clamped addresses + validity flags): the
        // kernel is latency-bound, branches around the loads would serialise them
        float v[9][V];
        bool ok[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int iw = ow * 2 - pl + kw;
                const int iwc = min(max(iw, 0), W - 1);
                ok[kh * 3 + kw] = okh[kh] && iw == iwc;
                load_vec(z + (((size_t)bi * H + ihc[kh]) * W + iwc) * C + c0, v[kh * 3 + kw]);
            }
        }
        float rv[V];
        if (res) load_vec(res + idx * V, rv);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            if (!ok[t]) continue;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float y = fmaf(v[t][i], av[i], bv[i]);
                if (y > best[i]) { best[i] = y; arg[i] = t; }
            }
        }
        if (res) {
            if (ra) {
                float rav[V], rbv[V];
                load_coef<V>(ra, c0, rav);
                load_coef<V>(rb, c0, rbv);
#pragma unroll
                for (int i = 0; i < V; ++i) best[i] += fmaf(rv[i], rav[i], rbv[i]);
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) best[i] += rv[i];
            }
        }
        store_vec(out + idx * V, best);
        if (argmax) {  // V codes in one store
            uint32_t w[V / 4];
#pragma unroll
            for (int i = 0; i < V / 4; ++i)
                w[i] = (uint32_t)arg[4 * i] | ((uint32_t)arg[4 * i + 1] << 8) | ((uint32_t)arg[4 * i + 2] << 16) |
                       ((uint32_t)arg[4 * i + 3] << 24);
            if (V == 8) *reinterpret_cast<uint2*>(argmax + idx * V) = make_uint2(w[0], w[V / 4 - 1]);
            else *reinterpret_cast<uint32_t*>(argmax + idx * V) = w[0];
        }
    }
}

// gin[b,h,w,c] = sum over the (<= 2x2) windows containing (h,w) whose argmax is (h,w)
template <typename T>
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const T* __restrict__ gout,
                                                          const unsigned char* __restrict__ argmax,
                                                          T* __restrict__ gin, int B, int H, int W, int C, int OH,
                                                          int OW, int pt, int pl) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int row = blockIdx.x;  // bi * H + h
    const int bi = row / H, h = row - bi * H;
    const int items = W * CV;
    // candidate window rows: kh with (h + pt - kh) even: kh = (h+pt)&1, and that + 2 (if <= 2) -- per CTA
    int ohc[2], khc[2];
    bool okh[2];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        const int kh = ((h + pt) & 1) + 2 * a;
        const int t = h + pt - kh;
        const int oh = t >> 1;
        okh[a] = kh <= 2 && t >= 0 && oh < OH;
        ohc[a] = okh[a] ? oh : 0;
        khc[a] = kh;
    }
    for (int it = blockIdx.y * blockDim.x + threadIdx.x; it < items; it += blockDim.x * gridDim.y) {
        const int w = it / CV, cv = it - w * CV;
        const int c0 = cv * V;
        float acc[V];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] = 0.f;
        // loads first (clamped, flagged), compares after
        float g[4][V];
        uint32_t am[4][2];
        int code[4];
        bool ok[4];
        int nc = 0;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const int kw = ((w + pl) & 1) + 2 * bb;
                const int u = w + pl - kw;
                const int ow = u >> 1;
                const bool okw = kw <= 2 && u >= 0 && ow < OW;
                ok[nc] = okh[a] && okw;
                code[nc] = khc[a] * 3 + kw;
                const size_t o = (((size_t)bi * OH + ohc[a]) * OW + (okw ? ow : 0)) * C + c0;
                load_vec(gout + o, g[nc]);
                if (V == 8) {
                    const uint2 q = *reinterpret_cast<const uint2*>(argmax + o);
                    am[nc][0] = q.x; am[nc][1] = q.y;
                } else {
                    am[nc][0] = *reinterpret_cast<const uint32_t*>(argmax + o); am[nc][1] = 0;
                }
                ++nc;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (!ok[k]) continue;
#pragma unroll
            for (int i = 0; i < V; ++i)
                if ((int)((am[k][i >> 2] >> (8 * (i & 3))) & 0xffu) == code[k]) acc[i] += g[k][i];
        }
        store_vec(gin + ((size_t)row * items + it) * V, acc);
    }
}

// out[b,oh,ow,:] = in[b,2oh+off_h,2ow+off_w,:]   (off 0: what a 1x1 stride-2 'same' convolution reads, and a 3x3
// stride-2 'same' convolution on an ODD-sized dimension = its stride-1 result at even positions; off 1: the
// same on an EVEN-sized dimension = the stride-1 result at odd positions, because TF pads only at the end)
template <typename T>
__global__ void __launch_bounds__(256) gather_s2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H,
                                                        int W, int C, int OH, int OW, int off_h, int off_w) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const long long n = (long long)B * OH * OW * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const unsigned uidx = (unsigned)idx;  // host guarantees < 2^31 items: 32-bit div/mod only
    const int cv = (int)(uidx % CV);
    unsigned r = uidx / CV;
    const int ow = (int)(r % OW);
    r /= OW;
    const int oh = (int)(r % OH);
    const int bi = (int)(r / OH);
    float v[V];
    load_vec(in + (((size_t)bi * H + 2 * oh + off_h) * W + 2 * ow + off_w) * C + cv * V, v);
    store_vec(out + idx * V, v);
}

// adjoint of gather_s2: out[b,h,w,:] = in[b,(h-off)/2,(w-off)/2,:] where both are integral, else 0
template <typename T>
__global__ void __launch_bounds__(256) scatter_s2_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int H,
                                                         int W, int C, int OH, int OW, int off_h, int off_w) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const long long n = (long long)B * H * W * CV;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const unsigned uidx = (unsigned)idx;
    const int cv = (int)(uidx % CV);
    unsigned r = uidx / CV;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int bi = (int)(r / H);
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = 0.f;
    const int th = h - off_h, tw = w - off_w;
    if (th >= 0 && tw >= 0 && !(th & 1) && !(tw & 1) && (th >> 1) < OH && (tw >> 1) < OW)
        load_vec(in + (((size_t)bi * OH + (th >> 1)) * OW + (tw >> 1)) * C + cv * V, v);
    store_vec(out + idx * V, v);
}

// grid of the row-per-CTA pooling kernels: x = rows, y = chunks of a row so that a thread handles ~2 items
dim3 row_grid(int rows, int items) {
    int gy = (items + 511) / 512;
    if (gy < 1) gy = 1;
    if (gy > 16) gy = 16;
    return dim3((unsigned)rows, (unsigned)gy);
}

int check_pool(const char* who, int dtype, int B, int H, int W, int C) {
    SPNET_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0, "%s: bad shape", who);
    SPNET_REQUIRE((long long)B * H * W * C < 0x7fffffffLL, "%s: tensor too large for 32-bit indexing", who);
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_REQUIRE(C % V == 0, "%s: C=%d must be a multiple of %d", who, C, V);
    return SPNET_OK;
}

}  // namespace

extern "C" {

int spnet_maxpool3s2_add_fwd(const void* z, const float* a, const float* b, const void* res, const float* ra,
                             const float* rb, void* out, unsigned char* argmax, int dtype, int B, int H, int W,
                             int C, cudaStream_t stream) {
    int rc = check_pool("maxpool3s2_add_fwd", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(z && out, "maxpool3s2_add_fwd: null pointer");
    SPNET_REQUIRE((a == nullptr) == (b == nullptr) && (ra == nullptr) == (rb == nullptr),
                  "maxpool3s2_add_fwd: affine parameters come in pairs");
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_DISPATCH_DTYPE(dtype, (maxpool_add_fwd_kernel<T><<<row_grid(B * OH, OW * (C / V)), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(z), a, b, reinterpret_cast<const T*>(res), ra, rb,
                                    reinterpret_cast<T*>(out), argmax, B, H, W, C, OH, OW, same_pad_before(H),
                                    same_pad_before(W))));
    return spnet_check_launch("maxpool3s2_add_fwd");
}

int spnet_maxpool3s2_bwd(const void* gout, const unsigned char* argmax, void* gin, int dtype, int B, int H, int W,
                         int C, cudaStream_t stream) {
    int rc = check_pool("maxpool3s2_bwd", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(gout && argmax && gin, "maxpool3s2_bwd: null pointer");
    const int OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_DISPATCH_DTYPE(dtype, (maxpool_bwd_kernel<T><<<row_grid(B * H, W * (C / V)), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(gout), argmax, reinterpret_cast<T*>(gin), B, H, W, C,
                                    OH, OW, same_pad_before(H), same_pad_before(W))));
    return spnet_check_launch("maxpool3s2_bwd");
}

// MaxPooling2D(3, strides 2, 'valid') (Inception-ResNet stem / reductions): the same kernels with no
// padding and OH = (H-3)/2+1. out [B,OH,OW,C]; argmax nullable (inference).
int spnet_maxpool3s2_valid_fwd(const void* z, void* out, unsigned char* argmax, int dtype, int B, int H, int W, int C,
                               cudaStream_t stream) {
    int rc = check_pool("maxpool3s2_valid_fwd", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(z && out && H >= 3 && W >= 3, "maxpool3s2_valid_fwd: bad args");
    const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_DISPATCH_DTYPE(dtype, (maxpool_add_fwd_kernel<T><<<row_grid(B * OH, OW * (C / V)), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(z), nullptr, nullptr, nullptr, nullptr, nullptr,
                                    reinterpret_cast<T*>(out), argmax, B, H, W, C, OH, OW, 0, 0)));
    return spnet_check_launch("maxpool3s2_valid_fwd");
}

int spnet_maxpool3s2_valid_bwd(const void* gout, const unsigned char* argmax, void* gin, int dtype, int B, int H, int W,
                               int C, cudaStream_t stream) {
    int rc = check_pool("maxpool3s2_valid_bwd", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(gout && argmax && gin && H >= 3 && W >= 3, "maxpool3s2_valid_bwd: bad args");
    const int OH = (H - 3) / 2 + 1, OW = (W - 3) / 2 + 1;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_DISPATCH_DTYPE(dtype, (maxpool_bwd_kernel<T><<<row_grid(B * H, W * (C / V)), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(gout), argmax, reinterpret_cast<T*>(gin), B, H, W, C, OH, OW,
                                    0, 0)));
    return spnet_check_launch("maxpool3s2_valid_bwd");
}

// per dimension: off = 0 takes positions 0,2,4,... (ceil(n/2) of them), off = 1 takes 1,3,5,... (n/2 of them)
int spnet_gather_s2(const void* in, void* out, int dtype, int B, int H, int W, int C, int off_h, int off_w,
                    cudaStream_t stream) {
    int rc = check_pool("gather_s2", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(in && out && (off_h == 0 || off_h == 1) && (off_w == 0 || off_w == 1), "gather_s2: bad args");
    const int OH = (H + 1 - off_h) / 2, OW = (W + 1 - off_w) / 2;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    const long long n = (long long)B * OH * OW * (C / V);
    SPNET_DISPATCH_DTYPE(dtype, (gather_s2_kernel<T><<<ceil_div(n, 256), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), B, H, W, C, OH, OW, off_h, off_w)));
    return spnet_check_launch("gather_s2");
}

// adjoint of spnet_gather_s2: in [B,OH,OW,C] (OH, OW as above) -> out [B,H,W,C], zero elsewhere
int spnet_scatter_s2(const void* in, void* out, int dtype, int B, int H, int W, int C, int off_h, int off_w,
                     cudaStream_t stream) {
    int rc = check_pool("scatter_s2", dtype, B, H, W, C);
    if (rc) return rc;
    SPNET_REQUIRE(in && out && (off_h == 0 || off_h == 1) && (off_w == 0 || off_w == 1), "scatter_s2: bad args");
    const int OH = (H + 1 - off_h) / 2, OW = (W + 1 - off_w) / 2;
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    const long long n = (long long)B * H * W * (C / V);
    SPNET_DISPATCH_DTYPE(dtype, (scatter_s2_kernel<T><<<ceil_div(n, 256), 256, 0, stream>>>(
                                    reinterpret_cast<const T*>(in), reinterpret_cast<T*>(out), B, H, W, C, OH, OW, off_h, off_w)));
    return spnet_check_launch("scatter_s2");
}

}  // extern "C"
