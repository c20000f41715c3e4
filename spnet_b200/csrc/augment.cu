// AugmentOnTheFly on the device (reference: spnet/callbacks.py:272-341 calling
// spnet/augmentation.py:117-135 cutout_inplace and :159-180 salt_n_pepa_inplace; blur_inplace :66-71 discards
// its result and is a no-op). The reference rewrites the training set on the host once per epoch with numpy's
// global RNG; here the pristine frames and the augmented copy both live in HBM, one CTA rewrites one frame,
// and the random draws come from a counter-based generator keyed by (seed, frame, draw), so an epoch's
// augmentation is reproducible and independent of launch geometry. The distributions are the reference's:
//   cutout : n ~ U{0..max_regions}; corner ~ (U{0..H-minsize-1}, U{0..W-minsize-1}); extents ~ U{minsize..maxsize-1};
//            the rectangle is clipped to [.., H-1) x [.., W-1); fill ~ U(min(frame), max(frame)); later rectangles
//            overwrite earlier ones
//   salt & pepper : with probability 1/2; ceil(amount*size*svp) salt points at max(frame), then
//            ceil(amount*size*(1-svp)) pepper points at min(frame) (min / max AFTER cutout), coordinates
//            ~ (U{0..H-2}, U{0..W-2}) as np.random.randint(0, dim - 1) draws them
#include "common.cuh"

namespace {

constexpr int kMaxRegions = 16;
constexpr int kThreads = 512;

// splitmix64 finaliser over a 3-word key
__device__ __forceinline__ uint64_t rnd64(uint64_t seed, uint64_t frame, uint64_t draw) {
    uint64_t z = seed + 0x9e3779b97f4a7c15ULL * (frame + 1) + 0xbf58476d1ce4e5b9ULL * (draw + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    z ^= z >> 31;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
// integer uniform on [lo, hi)
__device__ __forceinline__ int rnd_int(uint64_t r, int lo, int hi) {
    const uint32_t span = (uint32_t)(hi > lo ? hi - lo : 1);
    return lo + (int)(((r >> 32) * (uint64_t)span) >> 32);
}
__device__ __forceinline__ float rnd_unit(uint64_t r) { return (float)(r >> 40) * (1.0f / 16777216.0f); }

struct MinMax { float lo, hi; };
__device__ MinMax block_minmax(float lo, float hi, float* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();  // red may still be read from a previous call
    if (lane == 0) { red[warp] = lo; red[32 + warp] = hi; }
    __syncthreads();
    lo = red[0]; hi = red[32];
    for (int w = 1; w < kThreads / 32; ++w) { lo = fminf(lo, red[w]); hi = fmaxf(hi, red[32 + w]); }
    MinMax m = {lo, hi};
    return m;
}

__global__ void __launch_bounds__(kThreads) augment_kernel(const float* __restrict__ x_orig, float* __restrict__ x, int H,
                                                         int W, int C, unsigned long long seed, int max_regions,
                                                         int minsize, int maxsize, float sp_prob, float sp_amount,
                                                         float salt_vs_pepper) {
    __shared__ float red[64];
    __shared__ int rect[kMaxRegions][4];
    __shared__ float rval[kMaxRegions];
    __shared__ int s_nreg;
    const uint64_t frame = blockIdx.x;
    const long long npx = (long long)H * W, nel = npx * C;
    const float* src = x_orig + frame * nel;
    float* dst = x + frame * nel;
    // ---- pass 1: extrema of the pristine frame
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = threadIdx.x; i < nel; i += kThreads) {
        const float v = src[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
    const MinMax m0 = block_minmax(lo, hi, red);
    // ---- rectangles (draws 0 .. 5*max_regions)
    if (threadIdx.x == 0) {
        const int n = rnd_int(rnd64(seed, frame, 0), 0, max_regions + 1);
        s_nreg = n;
        for (int r = 0; r < n; ++r) {
            const int y0 = rnd_int(rnd64(seed, frame, 1 + 5 * r), 0, H - minsize);
            const int x0 = rnd_int(rnd64(seed, frame, 2 + 5 * r), 0, W - minsize);
            const int eh = rnd_int(rnd64(seed, frame, 3 + 5 * r), minsize, maxsize);
            const int ew = rnd_int(rnd64(seed, frame, 4 + 5 * r), minsize, maxsize);
            rect[r][0] = y0; rect[r][1] = min(y0 + eh, H - 1);
            rect[r][2] = x0; rect[r][3] = min(x0 + ew, W - 1);
            rval[r] = m0.lo + (m0.hi - m0.lo) * rnd_unit(rnd64(seed, frame, 5 + 5 * r));
        }
    }
    __syncthreads();
    const int nreg = s_nreg;
    // ---- pass 2: copy with the rectangles applied (the last one that covers a pixel wins), extrema of the result
    lo = INFINITY; hi = -INFINITY;
    for (long long p = threadIdx.x; p < npx; p += kThreads) {
        const int py = (int)(p / W), px = (int)(p - (long long)py * W);
        int hit = -1;
        for (int r = 0; r < nreg; ++r)
            if (py >= rect[r][0] && py < rect[r][1] && px >= rect[r][2] && px < rect[r][3]) hit = r;
        for (int c = 0; c < C; ++c) {
            const float v = hit >= 0 ? rval[hit] : src[p * C + c];
            dst[p * C + c] = v;
            lo = fminf(lo, v); hi = fmaxf(hi, v);
        }
    }
    const MinMax m1 = block_minmax(lo, hi, red);  // also orders pass 2's stores before the point writes below
    // ---- salt, then pepper
    const uint64_t d0 = 1 + 5 * (uint64_t)kMaxRegions;
    if (rnd_unit(rnd64(seed, frame, d0)) >= sp_prob) return;  // uniform over the block
    const int n_salt = (int)ceilf(sp_amount * (float)nel * salt_vs_pepper);
    const int n_pepper = (int)ceilf(sp_amount * (float)nel * (1.0f - salt_vs_pepper));
    for (int i = threadIdx.x; i < n_salt; i += kThreads) {
        const int py = rnd_int(rnd64(seed, frame, d0 + 1 + 2 * (uint64_t)i), 0, H - 1);
        const int px = rnd_int(rnd64(seed, frame, d0 + 2 + 2 * (uint64_t)i), 0, W - 1);
        for (int c = 0; c < C; ++c) dst[((long long)py * W + px) * C + c] = m1.hi;
    }
    __syncthreads();
    const uint64_t d1 = d0 + 1 + 2 * (uint64_t)n_salt;
    for (int i = threadIdx.x; i < n_pepper; i += kThreads) {
        const int py = rnd_int(rnd64(seed, frame, d1 + 2 * (uint64_t)i), 0, H - 1);
        const int px = rnd_int(rnd64(seed, frame, d1 + 1 + 2 * (uint64_t)i), 0, W - 1);
        for (int c = 0; c < C; ++c) dst[((long long)py * W + px) * C + c] = m1.lo;
    }
}

}  // namespace

extern "C" {

// x[n,H,W,C] (fp32) = augmented copy of x_orig; one launch per epoch. Defaults of the reference:
// max_regions 6, minsize 11, maxsize 75, sp_prob 0.5, sp_amount 0.004, salt_vs_pepper 0.2.
int spnet_augment_on_the_fly(const float* x_orig, float* x, int n, int H, int W, int C, long long seed,
                             int max_regions, int minsize, int maxsize, float sp_prob, float sp_amount,
                             float salt_vs_pepper, cudaStream_t stream) {
    SPNET_REQUIRE(x_orig && x && x_orig != x, "augment_on_the_fly: needs distinct source and destination");
    SPNET_REQUIRE(n > 0 && H > 1 && W > 1 && C > 0, "augment_on_the_fly: bad shape");
    SPNET_REQUIRE(max_regions >= 0 && max_regions <= kMaxRegions, "augment_on_the_fly: max_regions must be 0..%d", kMaxRegions);
    SPNET_REQUIRE(minsize > 0 && maxsize > minsize && H > minsize && W > minsize, "augment_on_the_fly: bad region sizes");
    augment_kernel<<<n, kThreads, 0, stream>>>(x_orig, x, H, W, C, (unsigned long long)seed, max_regions, minsize, maxsize, sp_prob, sp_amount,
                                               salt_vs_pepper);
    return spnet_check_launch("augment_on_the_fly");
}

}  // extern "C"
