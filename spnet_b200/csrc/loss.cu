// Fused YOLO-ellipse loss (value, 5-term breakdown and dL/dy_pred in one launch),
// selective sigmoid, and detection decode.
//
// Follows the reference's custom_loss (spnet/models.py:564-589) and its numpy
// twin my_loss (:594-633); constants from :557-562; column layout from
// spnet/config.py:30-38 (8 variables per predictor:
// cx, cy, a, b, cos2t, sin2t, noobj, rings).
#include "common.cuh"

namespace {

constexpr float kLambdaCenter = 2.0f;
constexpr float kLambdaSize = 1.0f;
constexpr float kLambdaAngle = 3.0f;
constexpr float kLambdaNoobj = 0.3f;
constexpr float kLambdaClass = 5.0f;

constexpr int kLossThreads = 1024;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// One warp per sample, lanes stride over predictors. Single CTA so that the
// batch reduction has a fixed order (bitwise reproducible run to run).
// out6 = [total, center, size, angle, noobj, class], each already divided by
// ncols and averaged over the batch, as my_loss returns them.
__global__ void __launch_bounds__(kLossThreads) yolo_ellipse_loss_kernel(
    const float* __restrict__ y_true, const float* __restrict__ y_pred, int batch, int ncols,
    int hybrid, int sel_sigmoid, float* __restrict__ out6, float* __restrict__ grad) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int npred = ncols >> 3;
    const float c = 1.0f / ((float)ncols * (float)batch);
    float acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // center,size,angle,noobj,class

    for (int s = warp; s < batch; s += nwarps) {
        const float* yt = y_true + (size_t)s * ncols;
        const float* yp = y_pred + (size_t)s * ncols;
        float* g = grad ? grad + (size_t)s * ncols : nullptr;
        for (int p = lane; p < npred; p += 32) {
            float t[8], q[8];
            const float4 t0 = *reinterpret_cast<const float4*>(yt + p * 8);
            const float4 t1 = *reinterpret_cast<const float4*>(yt + p * 8 + 4);
            const float4 q0 = *reinterpret_cast<const float4*>(yp + p * 8);
            const float4 q1 = *reinterpret_cast<const float4*>(yp + p * 8 + 4);
            t[0] = t0.x; t[1] = t0.y; t[2] = t0.z; t[3] = t0.w; t[4] = t1.x; t[5] = t1.y; t[6] = t1.z; t[7] = t1.w;
            q[0] = q0.x; q[1] = q0.y; q[2] = q0.z; q[3] = q0.w; q[4] = q1.x; q[5] = q1.y; q[6] = q1.z; q[7] = q1.w;

            float dsig = 1.0f;  // d(noobj activation)/d(raw)
            if (sel_sigmoid) {
                const float sg = sigmoidf_(q[6]);
                dsig = sg * (1.0f - sg);
                q[6] = sg;
            }
            const float pobj = 1.0f - t[6];
            const float ab = t[2] - t[3];
            const float ab2 = ab * ab;
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = q[i] - t[i];

            acc[0] += pobj * (e[0] * e[0]) + pobj * (e[1] * e[1]);
            acc[1] += pobj * (e[2] * e[2]) + pobj * (e[3] * e[3]);
            acc[2] += pobj * (e[4] * e[4]) * ab2 + pobj * (e[5] * e[5]) * ab2;
            acc[4] += pobj * (e[7] * e[7]);
            float gno;
            if (hybrid) {
                const float z = q[6];
                acc[3] += fmaxf(0.0f, z) - z * t[6] + log1pf(expf(-fabsf(z)));
                gno = kLambdaNoobj * (sigmoidf_(z) - t[6]);
            } else {
                acc[3] += e[6] * e[6];
                gno = 2.0f * kLambdaNoobj * e[6];
            }
            if (g) {
                float4 g0, g1;
                g0.x = 2.0f * kLambdaCenter * pobj * e[0] * c;
                g0.y = 2.0f * kLambdaCenter * pobj * e[1] * c;
                g0.z = 2.0f * kLambdaSize * pobj * e[2] * c;
                g0.w = 2.0f * kLambdaSize * pobj * e[3] * c;
                g1.x = 2.0f * kLambdaAngle * pobj * ab2 * e[4] * c;
                g1.y = 2.0f * kLambdaAngle * pobj * ab2 * e[5] * c;
                g1.z = gno * dsig * c;
                g1.w = 2.0f * kLambdaClass * pobj * e[7] * c;
                *reinterpret_cast<float4*>(g + p * 8) = g0;
                *reinterpret_cast<float4*>(g + p * 8 + 4) = g1;
            }
        }
    }
    __shared__ float part[32][5];
#pragma unroll
    for (int i = 0; i < 5; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) part[warp][i] = acc[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int w = 0; w < nwarps; ++w)
            for (int i = 0; i < 5; ++i) tot[i] += part[w][i];
        const float lam[5] = {kLambdaCenter, kLambdaSize, kLambdaAngle, kLambdaNoobj, kLambdaClass};
        float total = 0.f;
        for (int i = 0; i < 5; ++i) {
            const float v = lam[i] * tot[i] * c;
            out6[1 + i] = v;
            total += v;
        }
        out6[0] = total;
    }
}

// y[:, j] = sigmoid(x[:, j]) for j in range(start, end, skip), identity elsewhere
// (SelectiveSigmoid, spnet/models.py:277-298).
__global__ void selective_sigmoid_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                             long long n, int ncols, int start, int end, int skip) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = (int)(i % ncols);
    const bool sel = j >= start && j < end && ((j - start) % skip == 0);
    const float v = x[i];
    y[i] = sel ? sigmoidf_(v) : v;
}

// dx = dy * y(1-y) on the selected columns (y = forward output), dy elsewhere.
__global__ void selective_sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                             float* __restrict__ dx, long long n, int ncols, int start,
                                             int end, int skip) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int j = (int)(i % ncols);
    const bool sel = j >= start && j < end && ((j - start) % skip == 0);
    const float yy = y[i];
    dx[i] = sel ? dy[i] * yy * (1.0f - yy) : dy[i];
}

// Decode (denorm_Y + cleanup_antinode_vars' integer part + existence test;
// spnet/utils.py:186-188, :56-64, :109-118). One thread per predictor.
//   denorm  [n, ncols] f32 : y*ranges + means, two roundings exactly as numpy does
//   ints    [n, npred, 5] i32 : round-half-even of cx, cy, a, b, noobj
//   exists  [n, npred] u8  : noobj==0 && rings>0 && a>=0 && b>=0
__global__ void decode_detections_kernel(const float* __restrict__ y, const float* __restrict__ means,
                                         const float* __restrict__ ranges, int n, int ncols,
                                         float* __restrict__ denorm, int* __restrict__ ints,
                                         unsigned char* __restrict__ exists) {
    const int npred = ncols >> 3;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n * npred) return;
    const int p = (int)(i % npred);
    const size_t row = (size_t)(i / npred) * ncols + (size_t)p * 8;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        v[k] = __fadd_rn(__fmul_rn(y[row + k], ranges[p * 8 + k]), means[p * 8 + k]);
        denorm[row + k] = v[k];
    }
    const int cx = __float2int_rn(v[0]), cy = __float2int_rn(v[1]);
    const int a = __float2int_rn(v[2]), b = __float2int_rn(v[3]);
    const int noobj = __float2int_rn(v[6]);
    int* o = ints + i * 5;
    o[0] = cx; o[1] = cy; o[2] = a; o[3] = b; o[4] = noobj;
    exists[i] = (noobj == 0 && v[7] > 0.0f && a >= 0 && b >= 0) ? 1 : 0;
}

}  // namespace

extern "C" {

int spnet_yolo_ellipse_loss(const float* y_true, const float* y_pred, int batch, int ncols, int hybrid,
                            int sel_sigmoid, float* out6, float* grad, cudaStream_t stream) {
    SPNET_REQUIRE(y_true && y_pred && out6, "yolo_ellipse_loss: null pointer");
    SPNET_REQUIRE(batch > 0 && ncols > 0 && ncols % 8 == 0, "yolo_ellipse_loss: bad shape %d x %d", batch, ncols);
    yolo_ellipse_loss_kernel<<<1, kLossThreads, 0, stream>>>(y_true, y_pred, batch, ncols, hybrid, sel_sigmoid,
                                                             out6, grad);
    return spnet_check_launch("yolo_ellipse_loss");
}

int spnet_selective_sigmoid_fwd(const float* x, float* y, int rows, int ncols, int start, int end, int skip,
                                cudaStream_t stream) {
    SPNET_REQUIRE(x && y && rows > 0 && ncols > 0 && skip > 0, "selective_sigmoid_fwd: bad args");
    const long long n = (long long)rows * ncols;
    selective_sigmoid_fwd_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(x, y, n, ncols, start, end, skip);
    return spnet_check_launch("selective_sigmoid_fwd");
}

int spnet_selective_sigmoid_bwd(const float* y, const float* dy, float* dx, int rows, int ncols, int start,
                                int end, int skip, cudaStream_t stream) {
    SPNET_REQUIRE(y && dy && dx && rows > 0 && ncols > 0 && skip > 0, "selective_sigmoid_bwd: bad args");
    const long long n = (long long)rows * ncols;
    selective_sigmoid_bwd_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(y, dy, dx, n, ncols, start, end, skip);
    return spnet_check_launch("selective_sigmoid_bwd");
}

int spnet_decode_detections(const float* y, const float* means, const float* ranges, int n, int ncols,
                            float* denorm, int* ints, unsigned char* exists, cudaStream_t stream) {
    SPNET_REQUIRE(y && means && ranges && denorm && ints && exists, "decode_detections: null pointer");
    SPNET_REQUIRE(n > 0 && ncols > 0 && ncols % 8 == 0, "decode_detections: bad shape");
    const long long total = (long long)n * (ncols / 8);
    decode_detections_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(y, means, ranges, n, ncols, denorm, ints,
                                                                       exists);
    return spnet_check_launch("decode_detections");
}

}  // extern "C"
