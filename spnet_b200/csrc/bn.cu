// BatchNormalization (keras axis=-1, eps 1e-3, momentum 0.99; 43 instances in the
// reference graph: spnet/models.py:326-336 and keras.applications.Xception) split into
// the pieces a fused training step needs:
//   * batch statistics are accumulated (fp64 sum / sum of squares) by the PRODUCING
//     kernel's epilogue (GEMM, small conv); bn_finalize turns them into the per-channel
//     affine  y = a*z + b  that CONSUMERS apply on load, and updates the moving stats;
//   * bn_act / bn_add materialise BN(+activation / +residual) where a tensor has to exist;
//   * backward: bn_bwd_reduce -> bn_bwd_finalize -> bn_bwd_dz.
#include "common.cuh"

namespace {

__global__ void bn_finalize_kernel(double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, int unbiased,
                                   float* __restrict__ a, float* __restrict__ b, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = stats[c] / count;
    double var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[c] = 0.0;
    stats[C + c] = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float g = gamma ? gamma[c] : 1.0f;
    const float aa = g * rstd;
    a[c] = aa;
    b[c] = (beta ? beta[c] : 0.0f) - (float)mean * aa;
    save_mean[c] = (float)mean;
    save_rstd[c] = rstd;
    if (moving_mean) {
        const double uv = (unbiased && count > 1.0) ? var * count / (count - 1.0) : var;
        moving_mean[c] = momentum * moving_mean[c] + (1.0f - momentum) * (float)mean;
        moving_var[c] = momentum * moving_var[c] + (1.0f - momentum) * (float)uv;
    }
}

__global__ void bn_inference_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                           const float* __restrict__ mm, const float* __restrict__ mv, float eps,
                                           float* __restrict__ a, float* __restrict__ b, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float aa = (gamma ? gamma[c] : 1.0f) / sqrtf(mv[c] + eps);
    a[c] = aa;
    b[c] = (beta ? beta[c] : 0.0f) - mm[c] * aa;
}

// Channel-stationary mapping shared by the element-wise kernels below: a thread keeps one
// 16-byte channel vector (its per-channel coefficients stay in registers) and walks rows with a
// grid stride, UNROLL independent 16-byte loads in flight per tensor.
constexpr int UNROLL = 4;

// out = act(a*z + b) [+ x];  act: 0 none, 1 relu, 2 leaky-relu(0.1)
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* z, const float* __restrict__ a,
                                                       const float* __restrict__ b, int act, const T* x, T* out,
                                                       long long rows, int C, int cvb, int krows) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int cvl = threadIdx.x % cvb, rl = threadIdx.x / cvb;
    const int cv = blockIdx.y * cvb + cvl;
    if (cv >= CV) return;
    const int c0 = cv * V;
    float av[V], bv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { av[i] = a[c0 + i]; bv[i] = b[c0 + i]; }
    // each CTA walks a CONTIGUOUS range of rows (sequential DRAM pages), UNROLL*krows rows per iteration
    const long long per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const long long r_begin = (long long)blockIdx.x * per_cta;
    const long long r_end = r_begin + per_cta < rows ? r_begin + per_cta : rows;
    const long long stride = krows;
    for (long long r0 = r_begin + rl; r0 < r_end; r0 += stride * UNROLL) {
        float v[UNROLL][V], xr[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u * stride;
            if (r < r_end) {
                load_vec(z + r * C + c0, v[u]);
                if (x) load_vec(x + r * C + c0, xr[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u * stride;
            if (r < r_end) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    float y = fmaf(v[u][i], av[i], bv[i]);
                    if (act == 1) y = fmaxf(y, 0.f);
                    else if (act == 2) y = y > 0.f ? y : 0.1f * y;
                    if (x) y += xr[u][i];
                    v[u][i] = y;
                }
                store_vec(out + r * C + c0, v[u]);
            }
        }
    }
}

// stats[c] += sum g, stats[C+c] += sum g*xhat, xhat = (z-mean)*rstd.
// If relu_a is given, g is first masked in place by (relu_a*z+relu_b > 0) (act 1) or scaled
// by the leaky slope (act 2): the gradient of the activation that followed this BN.
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(T* __restrict__ g, const T* __restrict__ z,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            const float* __restrict__ relu_a,
                                                            const float* __restrict__ relu_b, int act,
                                                            double* __restrict__ stats, long long rows, int C,
                                                            int cvb, int krows) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int cvl = threadIdx.x % cvb, rl = threadIdx.x / cvb;
    const int cv = blockIdx.y * cvb + cvl;
    const bool active = cv < CV;
    const int c0 = cv * V;
    float s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    if (active) {
        float mu[V], rs[V], ra[V], rb[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            mu[i] = mean[c0 + i];
            rs[i] = rstd[c0 + i];
            ra[i] = relu_a ? relu_a[c0 + i] : 0.f;
            rb[i] = relu_a ? relu_b[c0 + i] : 0.f;
        }
        const long long per_cta = (rows + gridDim.x - 1) / gridDim.x;
        const long long r_begin = (long long)blockIdx.x * per_cta;
        const long long r_end = r_begin + per_cta < rows ? r_begin + per_cta : rows;
        const long long stride = krows;
        for (long long r0 = r_begin + rl; r0 < r_end; r0 += stride * UNROLL) {
            float gv[UNROLL][V], zv[UNROLL][V];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long r = r0 + u * stride;
                if (r < r_end) {
                    load_vec(g + r * C + c0, gv[u]);
                    load_vec(z + r * C + c0, zv[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const long long r = r0 + u * stride;
                if (r >= r_end) continue;
                if (relu_a) {
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float y = fmaf(zv[u][i], ra[i], rb[i]);
                        if (!(y > 0.f)) gv[u][i] = (act == 2) ? 0.1f * gv[u][i] : 0.f;
                    }
                    store_vec(g + r * C + c0, gv[u]);
                    // sums use the value as it will be re-read
#pragma unroll
                    for (int i = 0; i < V; ++i) gv[u][i] = round_to<T>(gv[u][i]);
                }
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    s1[i] += gv[u][i];
                    s2[i] = fmaf(gv[u][i], (zv[u][i] - mu[i]) * rs[i], s2[i]);
                }
            }
        }
    }
    __shared__ float red[2][256 * 8];
#pragma unroll
    for (int i = 0; i < V; ++i) { red[0][threadIdx.x * V + i] = s1[i]; red[1][threadIdx.x * V + i] = s2[i]; }
    __syncthreads();
    if (rl == 0 && active) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float t1 = 0.f, t2 = 0.f;
            for (int j = 0; j < krows; ++j) {
                t1 += red[0][(j * cvb + cvl) * V + i];
                t2 += red[1][(j * cvb + cvl) * V + i];
            }
            atomicAdd(stats + c0 + i, (double)t1);
            atomicAdd(stats + C + c0 + i, (double)t2);
        }
    }
}

__global__ void bn_bwd_finalize_kernel(double* __restrict__ stats, double count, float* __restrict__ dgamma,
                                       float* __restrict__ dbeta, float* __restrict__ c1, float* __restrict__ c2,
                                       int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double sg = stats[c], sgx = stats[C + c];
    stats[c] = 0.0;
    stats[C + c] = 0.0;
    if (dgamma) dgamma[c] = (float)sgx;
    if (dbeta) dbeta[c] = (float)sg;
    c1[c] = (float)(sg / count);
    c2[c] = (float)(sgx / count);
}

// out = a * (g - c1 - xhat*c2) = A*g + Bz*z + D with per-channel A = a, Bz = -a*rstd*c2,
// D = a*(rstd*c2*mean - c1)
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_dz_kernel(const T* g, const T* __restrict__ z,
                                                        const float* __restrict__ a,
                                                        const float* __restrict__ mean,
                                                        const float* __restrict__ rstd,
                                                        const float* __restrict__ c1,
                                                        const float* __restrict__ c2, T* out, long long rows, int C,
                                                        int cvb, int krows) {
    constexpr int V = VecN<T>::N;
    const int CV = C / V;
    const int cvl = threadIdx.x % cvb, rl = threadIdx.x / cvb;
    const int cv = blockIdx.y * cvb + cvl;
    if (cv >= CV) return;
    const int c0 = cv * V;
    float A[V], Bz[V], D[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const float aa = a[c0 + i], k = rstd[c0 + i] * c2[c0 + i];
        A[i] = aa;
        Bz[i] = -aa * k;
        D[i] = aa * (k * mean[c0 + i] - c1[c0 + i]);
    }
    // each CTA walks a CONTIGUOUS range of rows (sequential DRAM pages), UNROLL*krows rows per iteration
    const long long per_cta = (rows + gridDim.x - 1) / gridDim.x;
    const long long r_begin = (long long)blockIdx.x * per_cta;
    const long long r_end = r_begin + per_cta < rows ? r_begin + per_cta : rows;
    const long long stride = krows;
    for (long long r0 = r_begin + rl; r0 < r_end; r0 += stride * UNROLL) {
        float gv[UNROLL][V], zv[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u * stride;
            if (r < r_end) {
                load_vec(g + r * C + c0, gv[u]);
                load_vec(z + r * C + c0, zv[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u * stride;
            if (r < r_end) {
#pragma unroll
                for (int i = 0; i < V; ++i) gv[u][i] = fmaf(gv[u][i], A[i], fmaf(zv[u][i], Bz[i], D[i]));
                store_vec(out + r * C + c0, gv[u]);
            }
        }
    }
}

struct ChanGrid { int cvb, krows; dim3 grid; };
static ChanGrid chan_grid(int CV, long long rows, int waves) {
    ChanGrid c;
    const int nchunks = ceil_div(CV, 128);
    c.cvb = ceil_div(CV, nchunks);
    c.krows = 256 / c.cvb;
    if (c.krows < 1) c.krows = 1;
    long long gx = (rows + c.krows - 1) / c.krows;
    const long long cap = (148LL * waves + nchunks - 1) / nchunks;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    c.grid = dim3((unsigned)gx, nchunks);
    return c;
}

int check_rc(const char* who, int dtype, long long rows, int C) {
    SPNET_REQUIRE(rows > 0 && C > 0, "%s: bad shape", who);
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    SPNET_REQUIRE(C % V == 0, "%s: C=%d must be a multiple of %d", who, C, V);
    return SPNET_OK;
}

}  // namespace

extern "C" {

int spnet_bn_finalize(double* stats, long long count, const float* gamma, const float* beta, float eps,
                      float momentum, int unbiased_moving_var, float* a, float* b, float* save_mean,
                      float* save_rstd, float* moving_mean, float* moving_var, int C, cudaStream_t stream) {
    SPNET_REQUIRE(stats && a && b && save_mean && save_rstd && C > 0 && count > 0, "bn_finalize: bad args");
    SPNET_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize: moving stats come in pairs");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(stats, (double)count, gamma, beta, eps, momentum,
                                                             unbiased_moving_var, a, b, save_mean, save_rstd,
                                                             moving_mean, moving_var, C);
    return spnet_check_launch("bn_finalize");
}

int spnet_bn_inference_affine(const float* gamma, const float* beta, const float* moving_mean,
                              const float* moving_var, float eps, float* a, float* b, int C,
                              cudaStream_t stream) {
    SPNET_REQUIRE(moving_mean && moving_var && a && b && C > 0, "bn_inference_affine: bad args");
    bn_inference_affine_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(gamma, beta, moving_mean, moving_var, eps, a,
                                                                     b, C);
    return spnet_check_launch("bn_inference_affine");
}

// out = act(a*z+b) [+ x]   (x nullable; out may alias z or x)
int spnet_bn_apply(const void* z, const float* a, const float* b, int act, const void* x, void* out, int dtype,
                   long long rows, int C, cudaStream_t stream) {
    int rc = check_rc("bn_apply", dtype, rows, C);
    if (rc) return rc;
    SPNET_REQUIRE(z && a && b && out && act >= 0 && act <= 2, "bn_apply: bad args");
    const ChanGrid cg = chan_grid(C / (dtype == SPNET_BF16 ? 8 : 4), rows, 8);
    SPNET_DISPATCH_DTYPE(dtype, (bn_apply_kernel<T><<<cg.grid, cg.cvb * cg.krows, 0, stream>>>(
                                    reinterpret_cast<const T*>(z), a, b, act, reinterpret_cast<const T*>(x),
                                    reinterpret_cast<T*>(out), rows, C, cg.cvb, cg.krows)));
    return spnet_check_launch("bn_apply");
}

int spnet_bn_bwd_reduce(void* g, const void* z, const float* save_mean, const float* save_rstd,
                        const float* relu_a, const float* relu_b, int act, double* stats, int dtype,
                        long long rows, int C, cudaStream_t stream) {
    int rc = check_rc("bn_bwd_reduce", dtype, rows, C);
    if (rc) return rc;
    SPNET_REQUIRE(g && z && save_mean && save_rstd && stats, "bn_bwd_reduce: null pointer");
    SPNET_REQUIRE((relu_a == nullptr) == (relu_b == nullptr), "bn_bwd_reduce: mask affine comes in pairs");
    const int V = dtype == SPNET_BF16 ? 8 : 4;
    const int CV = C / V;
    const int nchunks = ceil_div(CV, 128);
    const int cvb = ceil_div(CV, nchunks);
    int krows = 256 / cvb;
    if (krows < 1) krows = 1;
    int gx = ceil_div(rows, krows);
    const int cap = (4 * 148 + nchunks - 1) / nchunks;
    if (gx > cap) gx = cap;
    dim3 grid(gx, nchunks);
    SPNET_DISPATCH_DTYPE(dtype, (bn_bwd_reduce_kernel<T><<<grid, cvb * krows, 0, stream>>>(
                                    reinterpret_cast<T*>(g), reinterpret_cast<const T*>(z), save_mean, save_rstd,
                                    relu_a, relu_b, act, stats, rows, C, cvb, krows)));
    return spnet_check_launch("bn_bwd_reduce");
}

int spnet_bn_bwd_finalize(double* stats, long long count, float* dgamma, float* dbeta, float* c1, float* c2,
                          int C, cudaStream_t stream) {
    SPNET_REQUIRE(stats && c1 && c2 && C > 0 && count > 0, "bn_bwd_finalize: bad args");
    bn_bwd_finalize_kernel<<<ceil_div(C, 128), 128, 0, stream>>>(stats, (double)count, dgamma, dbeta, c1, c2, C);
    return spnet_check_launch("bn_bwd_finalize");
}

// out = a*(g - c1 - xhat*c2)   (out may alias g)
int spnet_bn_bwd_dz(const void* g, const void* z, const float* a, const float* save_mean, const float* save_rstd,
                    const float* c1, const float* c2, void* out, int dtype, long long rows, int C,
                    cudaStream_t stream) {
    int rc = check_rc("bn_bwd_dz", dtype, rows, C);
    if (rc) return rc;
    SPNET_REQUIRE(g && z && a && save_mean && save_rstd && c1 && c2 && out, "bn_bwd_dz: null pointer");
    const ChanGrid cg = chan_grid(C / (dtype == SPNET_BF16 ? 8 : 4), rows, 8);
    SPNET_DISPATCH_DTYPE(dtype, (bn_bwd_dz_kernel<T><<<cg.grid, cg.cvb * cg.krows, 0, stream>>>(
                                    reinterpret_cast<const T*>(g), reinterpret_cast<const T*>(z), a, save_mean,
                                    save_rstd, c1, c2, reinterpret_cast<T*>(out), rows, C, cg.cvb, cg.krows)));
    return spnet_check_launch("bn_bwd_dz");
}

}  // extern "C"
