// Evaluation metrics on the device (reference: spnet/diagnostics.py).
//  * calc_errors (:13-60): ring-count / object-existence counters over every (image, predictor slot) and the
//    first predictor's centre error per image - exact (integer decisions on round-half-even, as Python's round).
//  * compute_iou (:85-120) for every (image, slot) in one launch: the reference draws two anti-aliased filled
//    ellipses with cv2 on a 512 x 384 canvas per pair and counts pixels on the CPU (calc_map does that for
//    2 x 72 x N masks x 10 thresholds). Here one CTA scans the pair's bounding box with the analytic ellipse
//    test, semi-axes enlarged by `margin` pixels to cover what cv2's anti-aliased edge touches (calibration and
//    tolerance: oracle/diagnostics_numpy.py). IoU = -1 where the reference returns -1.
#include "common.cuh"

namespace {

constexpr int VARS = 8;  // spnet/config.py:30

__global__ void __launch_bounds__(256) calc_errors_kernel(const float* __restrict__ yp, const float* __restrict__ yt,
                                                          int n, int ncols, int* __restrict__ counters,
                                                          float* __restrict__ pix_err) {
    const int slots = ncols / VARS;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    // counters: 0 ring_miscounts 1 ring_truecounts 2 total_obj 3 false_obj_pos 4 false_obj_neg 5 true_obj_pos 6 true_obj_neg
    int which = -1, ring = -1;
    if (idx < n * slots) {
        const int j = idx / slots, an = idx - j * slots;
        const float* p = yp + (size_t)j * ncols + an * VARS;
        const float* t = yt + (size_t)j * ncols + an * VARS;
        const int t_no = __float2int_rn(t[6]), p_no = __float2int_rn(p[6]);
        if (t_no == 0) {
            if (p_no == 0) { which = 5; ring = fabsf(__fsub_rn(t[7], p[7])) > 0.5f ? 0 : 1; }
            else which = 4;
        } else {
            which = p_no == 0 ? 3 : 6;
        }
        if (an == 0) {
            const float d0 = __fsub_rn(p[0], t[0]), d1 = __fsub_rn(p[1], t[1]);
            pix_err[j] = __fsqrt_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)));
        }
    }
    // warp-aggregated atomics
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const bool hit = which == c || ring == c || (c == 2 && (which == 4 || which == 5));
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if ((threadIdx.x & 31) == 0 && m) atomicAdd(counters + c, __popc(m));
    }
}

struct Ell { float cx, cy, ia, ib, ct, st; bool drawn; int x0, x1, y0, y1; };

__device__ __forceinline__ Ell make_ell(const float* a, float margin, int nx, int ny) {
    Ell e;
    e.drawn = a[6] < 0.5f;
    e.cx = a[0]; e.cy = a[1];
    const float sa = a[2] + margin, sb = a[3] + margin;
    e.ia = 1.0f / sa; e.ib = 1.0f / sb;
    const float th = -0.5f * atan2f(a[5], a[4]);
    e.ct = cosf(th); e.st = sinf(th);
    const float r = fmaxf(fabsf(sa), fabsf(sb)) + 1.0f;
    e.x0 = max(0, (int)floorf(e.cx - r)); e.x1 = min(nx - 1, (int)ceilf(e.cx + r));
    e.y0 = max(0, (int)floorf(e.cy - r)); e.y1 = min(ny - 1, (int)ceilf(e.cy + r));
    return e;
}
__device__ __forceinline__ bool inside(const Ell& e, float x, float y) {
    const float dx = x - e.cx, dy = y - e.cy;
    const float u = (dx * e.ct + dy * e.st) * e.ia, v = (dy * e.ct - dx * e.st) * e.ib;
    return u * u + v * v <= 1.0f;
}

__global__ void __launch_bounds__(256) ellipse_iou_kernel(const float* __restrict__ yp, const float* __restrict__ yt,
                                                          int ncols, int nx, int ny, float margin,
                                                          float* __restrict__ iou, int* __restrict__ counts) {
    const int slots = ncols / VARS;
    const int pair = blockIdx.x, j = pair / slots, an = pair - j * slots;
    const float* p = yp + (size_t)j * ncols + an * VARS;
    const float* t = yt + (size_t)j * ncols + an * VARS;
    __shared__ int red[2][8];
    if (t[6] > 0.99f) {  // empty true slot: the reference skips the pair
        if (threadIdx.x == 0) {
            iou[pair] = -1.0f;
            if (counts) { counts[2 * pair] = 0; counts[2 * pair + 1] = 0; }
        }
        return;
    }
    const Ell ep = make_ell(p, margin, nx, ny), et = make_ell(t, margin, nx, ny);
    int x0 = nx, x1 = -1, y0 = ny, y1 = -1;
    if (ep.drawn) { x0 = min(x0, ep.x0); x1 = max(x1, ep.x1); y0 = min(y0, ep.y0); y1 = max(y1, ep.y1); }
    if (et.drawn) { x0 = min(x0, et.x0); x1 = max(x1, et.x1); y0 = min(y0, et.y0); y1 = max(y1, et.y1); }
    int ni = 0, nu = 0;
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    if (bw > 0 && bh > 0) {
        const int npx = bw * bh;
        for (int i = threadIdx.x; i < npx; i += blockDim.x) {
            const int yy = i / bw, xx = i - yy * bw;
            const float x = (float)(x0 + xx), y = (float)(y0 + yy);
            const bool a = ep.drawn && inside(ep, x, y), b = et.drawn && inside(et, x, y);
            ni += a && b;
            nu += a || b;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
        nu += __shfl_xor_sync(0xffffffffu, nu, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ni; red[1][threadIdx.x >> 5] = nu; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int si = 0, su = 0;
        for (int w = 0; w < 8; ++w) { si += red[0][w]; su += red[1][w]; }
        iou[pair] = su > 0 ? (float)si / (float)su : -1.0f;  // nothing drawn at all: -1 like the reference
        if (counts) { counts[2 * pair] = si; counts[2 * pair + 1] = su; }
    }
}

}  // namespace

extern "C" {

// counters: int32 [7] (zeroed by the caller): ring_miscounts, ring_truecounts, total_obj, false_obj_pos,
// false_obj_neg, true_obj_pos, true_obj_neg; pix_err: fp32 [n]. yp / yt: denormalised [n, ncols] fp32.
int spnet_calc_errors(const float* yp, const float* yt, int n, int ncols, int* counters, float* pix_err,
                      cudaStream_t stream) {
    SPNET_REQUIRE(yp && yt && counters && pix_err && n > 0 && ncols > 0 && ncols % VARS == 0, "calc_errors: bad args");
    const int items = n * (ncols / VARS);
    calc_errors_kernel<<<ceil_div(items, 256), 256, 0, stream>>>(yp, yt, n, ncols, counters, pix_err);
    return spnet_check_launch("calc_errors");
}

// iou: fp32 [n, ncols/8]; counts: nullable int32 [n, ncols/8, 2] (intersection, union pixel counts)
int spnet_ellipse_iou(const float* yp, const float* yt, int n, int ncols, int nx, int ny, float margin, float* iou,
                      int* counts, cudaStream_t stream) {
    SPNET_REQUIRE(yp && yt && iou && n > 0 && ncols > 0 && ncols % VARS == 0 && nx > 0 && ny > 0, "ellipse_iou: bad args");
    ellipse_iou_kernel<<<n * (ncols / VARS), 256, 0, stream>>>(yp, yt, ncols, nx, ny, margin, iou, counts);
    return spnet_check_launch("ellipse_iou");
}

}  // extern "C"
