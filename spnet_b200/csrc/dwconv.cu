// Depthwise 3x3, stride 1, 'same' padding, NHWC — stand-alone data-gradient and
// weight-gradient entry points (the forward and the fused backward the engine runs are in
// dwconv_packed.cu). This is the depthwise half of keras SeparableConv2D as used by
// keras.applications.Xception (reference call site spnet/models.py:359).
//
// HBM-bound: every thread owns one 16-byte channel vector (8 bf16 / 4 fp32) of one
// pixel column and slides down a strip of rows, so each input vector is loaded
// from global memory once per (column, +-1 neighbour) and all arithmetic is fp32
// in registers. Xception is pre-activation (ReLU -> sepconv -> BN), and in
// training the producer's BatchNorm can only be applied once its batch statistics
// exist, so BN-apply (per-channel affine) + ReLU are fused on the LOAD side here;
// zero padding is applied after that transform, as TF does.
#include "tma.cuh"

namespace {

struct DwEpilogue {
    // All optional (nullptr = off). Used by the data-gradient pass.
    const void* mask_src;    // [B,H,W,C]  multiply result by (mask_a*src+mask_b > 0)
    const float* mask_a;     // [C] or nullptr (=1)
    const float* mask_b;     // [C] or nullptr (=0)
    const void* add_src;     // [B,H,W,C]  added after masking
    const void* add_strided; // [B,ceil(H/2),ceil(W/2),C] added at even (h,w) after masking
};

// ---- 4-channel group access (8 B for bf16, 16 B for fp32) --------------------------------------
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
    const uint2 r = *reinterpret_cast<const uint2*>(p);
    v[0] = __uint_as_float(r.x << 16); v[1] = __uint_as_float(r.x & 0xffff0000u);
    v[2] = __uint_as_float(r.y << 16); v[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
}

// raw (not yet widened) 4-channel group: lets a global load be issued early and consumed late —
// the first instruction that touches the loaded register is where an in-order warp stalls
template <typename T> struct Raw4;
template <> struct Raw4<float> { float4 r; };
template <> struct Raw4<bf16> { uint2 r; };
__device__ __forceinline__ Raw4<float> load_raw4(const float* p) { Raw4<float> x; x.r = *reinterpret_cast<const float4*>(p); return x; }
__device__ __forceinline__ Raw4<bf16> load_raw4(const bf16* p) { Raw4<bf16> x; x.r = *reinterpret_cast<const uint2*>(p); return x; }
__device__ __forceinline__ void unpack4(const Raw4<float>& x, float (&v)[4]) { v[0] = x.r.x; v[1] = x.r.y; v[2] = x.r.z; v[3] = x.r.w; }
__device__ __forceinline__ void unpack4(const Raw4<bf16>& x, float (&v)[4]) {
    v[0] = __uint_as_float(x.r.x << 16); v[1] = __uint_as_float(x.r.x & 0xffff0000u);
    v[2] = __uint_as_float(x.r.y << 16); v[3] = __uint_as_float(x.r.y & 0xffff0000u);
}

constexpr int DW_CB = 64;   // channels per CTA (one TMA box is CB channels wide)
constexpr int DW_G = 4;     // channels per thread

// ---- shared-memory reads through 32-bit shared-window addresses (LDS, no generic-address math) ----
__device__ __forceinline__ void lds4(uint32_t addr, float (&v)[4], const float*) {
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(addr));
}
__device__ __forceinline__ void lds4(uint32_t addr, float (&v)[4], const bf16*) {
    uint32_t a, b;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
    v[0] = __uint_as_float(a << 16); v[1] = __uint_as_float(a & 0xffff0000u);
    v[2] = __uint_as_float(b << 16); v[3] = __uint_as_float(b & 0xffff0000u);
}

// One gradient/input row (4 staged columns x 4 channels) into three running accumulators:
// P completes with tap row 2, Q gets tap row 1, R starts with tap row 0.
__device__ __forceinline__ void dw_row_fma(const float (&x)[4][DW_G], const float (&wt)[9][DW_G], float (&P)[2][DW_G],
                                           float (&Q)[2][DW_G], float (&R)[2][DW_G]) {
#pragma unroll
    for (int oc = 0; oc < 2; ++oc)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int i = 0; i < DW_G; ++i) {
                P[oc][i] = fmaf(x[oc + kw][i], wt[6 + kw][i], P[oc][i]);
                Q[oc][i] = fmaf(x[oc + kw][i], wt[3 + kw][i], Q[oc][i]);
                R[oc][i] = fmaf(x[oc + kw][i], wt[0 + kw][i], R[oc][i]);
            }
}

struct DwTileCoord { int b, h0, w0; };
__device__ __forceinline__ DwTileCoord dw_tile_coord(int tile, int tiles_w, int tiles_h, int TH, int TW) {
    DwTileCoord t;
    const int tw = tile % tiles_w;
    const int q = tile / tiles_w;
    t.w0 = tw * TW;
    t.h0 = (q % tiles_h) * TH;
    t.b = q / tiles_h;
    return t;
}

// Forward / data-gradient kernel. A persistent CTA owns one 64-channel chunk (blockIdx.y) and
// walks (image, row-tile, col-tile) tiles; each input tile INCLUDING its 1-pixel halo is fetched
// by ONE TMA box load ({64 ch, TW+2, TH+2, 1}; out-of-image halo is zero-filled by the TMA unit)
// into a double-buffered shared-memory stage, so the next tile streams in while the current one
// is computed. Thread = (4-channel group, 2 adjacent output columns); it slides down the rows
// with three running accumulators per column (rotated by a 3x unrolled loop, no register moves),
// reading each staged input vector once per row. EPI enables the optional mask / residual-add epilogue.
template <typename T, bool AFFINE, bool RELU, bool FLIP, bool EPI>
__global__ void __launch_bounds__(256, 2) dw3x3_tma_kernel(const __grid_constant__ CUtensorMap tm_in,
                                                           const float* __restrict__ k,
                                                           const float* __restrict__ in_a,
                                                           const float* __restrict__ in_b, T* __restrict__ out, int B,
                                                           int H, int W, int C, int TH, int TW, int tiles_h,
                                                           int tiles_w, DwEpilogue ep) {
    constexpr int G = DW_G, CB = DW_CB;
    constexpr uint32_t ES = sizeof(T);
    extern __shared__ uint8_t dw_smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    const uint32_t sbase = (smem_u32(dw_smem_raw) + 127u) & ~127u;
    const int cg = threadIdx.x & 15, colg = threadIdx.x >> 4;
    const int cbase = blockIdx.y * CB;
    const int c0 = cbase + cg * G;
    const bool c_ok = c0 < C;
    const int TWH = TW + 2;
    const uint32_t tile_bytes = (uint32_t)((TH + 2) * TWH * CB) * ES;
    const uint32_t row_bytes = (uint32_t)(TWH * CB) * ES;
    const int n_tiles = B * tiles_h * tiles_w;
    const int col0 = colg * 2;

    float wt[9][G], av[G], bv[G];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int ts = FLIP ? 8 - t : t;
#pragma unroll
        for (int i = 0; i < G; ++i) wt[t][i] = c_ok ? k[ts * C + c0 + i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < G; ++i) {
        av[i] = (AFFINE && c_ok) ? in_a[c0 + i] : 1.f;
        bv[i] = (AFFINE && c_ok) ? in_b[c0 + i] : 0.f;
    }
    const uint32_t bar0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0 && (int)blockIdx.x < n_tiles) {
        const DwTileCoord t = dw_tile_coord(blockIdx.x, tiles_w, tiles_h, TH, TW);
        mbar_expect_tx(bar0, tile_bytes);
        tma_load_4d(sbase, &tm_in, bar0, cbase, t.w0 - 1, t.h0 - 1, t.b);
    }
    const size_t rowC = (size_t)W * C;

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int cur = it & 1;
        const int nxt = tile + gridDim.x;
        if (threadIdx.x == 0 && nxt < n_tiles) {
            const DwTileCoord t = dw_tile_coord(nxt, tiles_w, tiles_h, TH, TW);
            mbar_expect_tx(bar0 + 8 * (cur ^ 1), tile_bytes);
            tma_load_4d(sbase + (cur ^ 1) * tile_bytes, &tm_in, bar0 + 8 * (cur ^ 1), cbase, t.w0 - 1, t.h0 - 1, t.b);
        }
        const DwTileCoord tc = dw_tile_coord(tile, tiles_w, tiles_h, TH, TW);
        const int h0 = tc.h0, w0 = tc.w0;
        const int rows = min(TH, H - h0) + 2;
        bool colok[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int iw = w0 - 1 + col0 + j;
            colok[j] = iw >= 0 && iw < W;
        }
        const bool st0 = c_ok && (w0 + col0) < W, st1 = c_ok && (w0 + col0 + 1) < W;
        uint32_t s_row = sbase + cur * tile_bytes + (uint32_t)(col0 * CB + cg * G) * ES;
        T* o_ptr = out + ((size_t)tc.b * H + h0) * rowC + (size_t)(w0 + col0) * C + c0;
        float A_[2][G], B_[2][G], C_[2][G];
#pragma unroll
        for (int oc = 0; oc < 2; ++oc)
#pragma unroll
            for (int i = 0; i < G; ++i) { A_[oc][i] = 0.f; B_[oc][i] = 0.f; C_[oc][i] = 0.f; }
        mbar_wait(bar0 + 8 * cur, (uint32_t)(it >> 1) & 1u);

        int r = 0;
        auto step = [&](float (&P)[2][G], float (&Q)[2][G], float (&R)[2][G]) {
            const int ih = h0 - 1 + r;
            const bool rowok = ih >= 0 && ih < H;
            float x[4][G];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                lds4(s_row + j * (CB * ES), x[j], (const T*)nullptr);
                if (AFFINE || RELU) {
                    const bool ok = rowok && colok[j];
#pragma unroll
                    for (int i = 0; i < G; ++i) {
                        float v = x[j][i];
                        if (AFFINE) v = fmaf(v, av[i], bv[i]);
                        if (RELU) v = fmaxf(v, 0.f);
                        x[j][i] = (AFFINE && !ok) ? 0.f : v;  // zero padding applies AFTER the BN/ReLU transform
                    }
                }
            }
            dw_row_fma(x, wt, P, Q, R);
            if (r >= 2) {  // P now holds output row h0 + r - 2
                if (EPI) {
                    const int oh = h0 + r - 2;
#pragma unroll
                    for (int oc = 0; oc < 2; ++oc) {
                        if (!(oc ? st1 : st0)) continue;
                        const int ow = w0 + col0 + oc;
                        const size_t o = (size_t)(o_ptr - out) + (size_t)oc * C;
                        if (ep.mask_src) {
                            float m[G];
                            load4(reinterpret_cast<const T*>(ep.mask_src) + o, m);
#pragma unroll
                            for (int i = 0; i < G; ++i) {
                                float v = m[i];
                                if (ep.mask_a) v = fmaf(v, ep.mask_a[c0 + i], ep.mask_b[c0 + i]);
                                if (!(v > 0.f)) P[oc][i] = 0.f;
                            }
                        }
                        if (ep.add_src) {
                            float m[G];
                            load4(reinterpret_cast<const T*>(ep.add_src) + o, m);
#pragma unroll
                            for (int i = 0; i < G; ++i) P[oc][i] += m[i];
                        }
                        if (ep.add_strided && ((oh | ow) & 1) == 0) {
                            const int H2 = (H + 1) >> 1, W2 = (W + 1) >> 1;
                            float m[G];
                            load4(reinterpret_cast<const T*>(ep.add_strided) +
                                      (((size_t)tc.b * H2 + (oh >> 1)) * W2 + (ow >> 1)) * C + c0, m);
#pragma unroll
                            for (int i = 0; i < G; ++i) P[oc][i] += m[i];
                        }
                    }
                }
                if (st0) store4(o_ptr, P[0]);
                if (st1) store4(o_ptr + C, P[1]);
                o_ptr += rowC;
            }
#pragma unroll
            for (int oc = 0; oc < 2; ++oc)
#pragma unroll
                for (int i = 0; i < G; ++i) P[oc][i] = 0.f;
            s_row += row_bytes;
            ++r;
        };
        while (r + 3 <= rows) { step(A_, B_, C_); step(B_, C_, A_); step(C_, A_, B_); }
        if (r < rows) { step(A_, B_, C_); if (r < rows) step(B_, C_, A_); }
        __syncthreads();  // every thread is done with buffer `cur` before it is refilled
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

static int make_nhwc_map(CUtensorMap* map, const void* ptr, int dtype, int B, int H, int W, int C, int box_w, int box_h) {
    PFN_cuTensorMapEncodeTiled enc = spnet_get_tensormap_encoder();
    if (!enc) {
        spnet_set_error("dwconv3x3: cuTensorMapEncodeTiled entry point not available");
        return SPNET_ERR_CUDA;
    }
    const cuuint64_t es = dtype == SPNET_BF16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {(cuuint32_t)DW_CB, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, dtype == SPNET_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        spnet_set_error("dwconv3x3: cuTensorMapEncodeTiled failed (%d) for [%d,%d,%d,%d]", (int)r, B, H, W, C);
        return SPNET_ERR_CUDA;
    }
    return SPNET_OK;
}

struct DwTiling { int TH, TW, tiles_h, tiles_w, threads, chunks, grid_x; size_t smem; };
static DwTiling dw_tiling(int dtype, int B, int H, int W, int C) {
    DwTiling t;
    t.TW = W > 16 ? 32 : (W > 8 ? 16 : 8);
    t.TH = H <= 12 ? H : 8;
    t.tiles_h = ceil_div(H, t.TH);
    t.tiles_w = ceil_div(W, t.TW);
    t.threads = 16 * (t.TW / 2);
    t.chunks = ceil_div(C, DW_CB);
    const size_t es = dtype == SPNET_BF16 ? 2 : 4;
    t.smem = 2 * (size_t)(t.TH + 2) * (t.TW + 2) * DW_CB * es + 128;
    int per_sm = 768 / t.threads;
    const int by_smem = (int)((220 * 1024) / t.smem);
    if (per_sm > by_smem) per_sm = by_smem;
    if (per_sm < 1) per_sm = 1;
    const long long n_tiles = (long long)B * t.tiles_h * t.tiles_w;
    long long gx = (148LL * per_sm + t.chunks - 1) / t.chunks;
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    t.grid_x = (int)gx;
    return t;
}

template <typename T, bool AF, bool RL, bool FL, bool EP>
int launch_dw_inst(const CUtensorMap& tm, const float* k, const float* in_a, const float* in_b, T* y, int B, int H,
                   int W, int C, const DwTiling& t, DwEpilogue ep, cudaStream_t stream) {
    auto kern = dw3x3_tma_kernel<T, AF, RL, FL, EP>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            spnet_set_error("dwconv3x3: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return SPNET_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid(t.grid_x, t.chunks);
    kern<<<grid, t.threads, t.smem, stream>>>(tm, k, in_a, in_b, y, B, H, W, C, t.TH, t.TW, t.tiles_h, t.tiles_w, ep);
    return spnet_check_launch("dw3x3");
}

template <typename T>
int launch_dw(const void* in, const float* k, const float* in_a, const float* in_b, int relu, int flip,
              void* out, int dtype, int B, int H, int W, int C, DwEpilogue ep, cudaStream_t stream) {
    const DwTiling t = dw_tiling(dtype, B, H, W, C);
    SPNET_REQUIRE(t.smem <= 200 * 1024, "dwconv3x3: tile does not fit shared memory");
    CUtensorMap tm;
    int rc = make_nhwc_map(&tm, in, dtype, B, H, W, C, t.TW + 2, t.TH + 2);
    if (rc) return rc;
    T* y = reinterpret_cast<T*>(out);
    // only the stand-alone data-gradient entry point still runs on this kernel (forward and the
    // fused backward live in dwconv_packed.cu)
    SPNET_REQUIRE(flip && !in_a && !relu, "dwconv3x3: internal: unexpected variant");
    return launch_dw_inst<T, false, false, true, true>(tm, k, in_a, in_b, y, B, H, W, C, t, ep, stream);
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
    SPNET_DISPATCH_DTYPE(dtype, return launch_dw<T>(gout, k, nullptr, nullptr, 0, 1, gin, dtype, B, H, W, C, ep, stream));
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
