// How fast can one SM land TMA tiles from L2? 1 CTA per SM, no consumer: thread(s) keep S boxes in flight and
// re-issue as each lands. Prints bytes/clk/SM and clk per box row for several box shapes.
// READ THE RESULT WITH CARE: one issuing thread shows ~420 cycles per box whatever the box size and two issuing
// warps show half of that - this is the cost of THIS LOOP's single-thread instruction stream (integer
// divisions, an ELECT / R2UR.BROADCAST sequence around every UTMALDG because the operands live in vector
// registers), not a limit of the TMA unit: with four issuers the SM lands ~74 B/clk. The same effect bounded the
// narrow GEMM / convolution tiles until the issue threads moved to the uniform datapath.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_rate tma_rate.cu
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

// issuers: number of warps whose lane 0 issues (each owns stages s % issuers == its index)
__global__ void __launch_bounds__(128, 1) rate_kernel(const __grid_constant__ CUtensorMap tm, int box_bytes, int box_rows,
                                                     int stages, int issuers, int n_boxes, int rows_per_cta, int inner_steps,
                                                     long long* cycles, int G, int prefetch) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[32];
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (prefetch && threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    const long long t0 = clock64();
    if (G > 1) {
        // one thread, G boxes per stage issued back to back (one wait + one expect_tx per G boxes)
        if (threadIdx.x == 0) {
            const int groups = n_boxes / G, boxes_per_pass = rows_per_cta / box_rows;
            const int row_base = blockIdx.x * rows_per_cta;
            for (int i = 0; i < groups; ++i) {
                const int s = i % stages, use = i / stages;
                if (use > 0) mbar_wait(smem_u32(&bars[s]), (use - 1) & 1);
                mbar_expect_tx(smem_u32(&bars[s]), box_bytes * G);
                for (int g = 0; g < G; ++g) {
                    const int b = (i * G + g) % (boxes_per_pass * inner_steps);
                    tma_load_2d(smem_u32(smem + (size_t)(s * G + g) * box_bytes), &tm, smem_u32(&bars[s]),
                                (b % inner_steps) * (box_bytes / box_rows / 2), row_base + (b / inner_steps) * box_rows);
                }
            }
            for (int s = 0; s < stages && s < groups; ++s) mbar_wait(smem_u32(&bars[s]), ((groups - 1 - s) / stages) & 1);
        }
    } else
    if (warp < issuers && lane == 0) {
        const int row_base = blockIdx.x * rows_per_cta;
        const int boxes_per_pass = rows_per_cta / box_rows;
        for (int i = warp; i < n_boxes; i += issuers) {
            const int s = i % stages;
            const int use = i / stages;
            if (use > 0) mbar_wait(smem_u32(&bars[s]), (use - 1) & 1);
            mbar_expect_tx(smem_u32(&bars[s]), box_bytes);
            const int b = i % (boxes_per_pass * inner_steps);
            tma_load_2d(smem_u32(smem + (size_t)s * box_bytes), &tm, smem_u32(&bars[s]), (b % inner_steps) * (box_bytes / box_rows / 2),
                        row_base + (b / inner_steps) * box_rows);
        }
        // drain
        for (int s = warp; s < stages && s < n_boxes; s += issuers) {
            const int last_use = (n_boxes - 1 - s) / stages;
            mbar_wait(smem_u32(&bars[s]), last_use & 1);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

int main() {
    PFN_cuTensorMapEncodeTiled enc = nullptr;
    {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
        enc = (PFN_cuTensorMapEncodeTiled)p;
    }
    const int n_cta = 148, rows_per_cta = 1024, cols = 256;  // 148 * 1024 rows x 256 bf16 = 77.6 MB: L2 resident mostly
    const long long rows = (long long)n_cta * rows_per_cta;
    void* d;
    cudaMalloc(&d, rows * cols * 2);
    cudaMemset(d, 0, rows * cols * 2);
    long long* dcyc;
    cudaMalloc(&dcyc, n_cta * sizeof(long long));
    struct V { const char* name; int inner; int brows; CUtensorMapSwizzle sw; int stages; int issuers; int G; int prefetch; };
    V vs[] = {
        {"64x128 S8 1 issuer", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 8, 1, 1, 0},
        {"64x128 S8 1 issuer prefetch", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 8, 1, 1, 1},
        {"64x128 S4 G2 (2 back to back)", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 4, 1, 2, 0},
        {"64x64  S4 G4 (4 back to back)", 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 4, 1, 4, 0},
        {"64x128 S2 G4 (4 back to back)", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 2, 1, 4, 0},
        {"64x128 S8 2 issuers", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 8, 2, 1, 0},
        {"64x128 S8 3 issuers", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 8, 3, 1, 0},
        {"64x64  S8 4 issuers", 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 8, 4, 1, 0},
        {"64x64  S12 4 issuers", 64, 64, CU_TENSOR_MAP_SWIZZLE_128B, 12, 4, 1, 0},
        {"64x32  S12 4 issuers", 64, 32, CU_TENSOR_MAP_SWIZZLE_128B, 12, 4, 1, 0},
        {"64x128 S2 1 issuer", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 2, 1, 1, 0},
        {"64x128 S1 1 issuer", 64, 128, CU_TENSOR_MAP_SWIZZLE_128B, 1, 1, 1, 0},
    };
    for (const V& v : vs) {
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * 2};
        cuuint32_t box[2] = {(cuuint32_t)v.inner, (cuuint32_t)v.brows}, estr[2] = {1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, v.sw,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-28s encode failed %d\n", v.name, (int)r); continue; }
        const int box_bytes = v.inner * v.brows * 2;
        const int smem = v.stages * v.G * box_bytes + 1024;
        cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        const int n_boxes = (8 << 20) / box_bytes;  // 8 MB per SM
        const int inner_steps = cols / v.inner;
        for (int rep = 0; rep < 2; ++rep)
            rate_kernel<<<n_cta, 128, smem>>>(tm, box_bytes, v.brows, v.stages, v.issuers, n_boxes, rows_per_cta, inner_steps, dcyc, v.G, v.prefetch);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%-28s failed: %s\n", v.name, cudaGetErrorString(e)); return 1; }
        std::vector<long long> h(n_cta);
        cudaMemcpy(h.data(), dcyc, n_cta * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0; for (long long c : h) avg += (double)c; avg /= n_cta;
        const double bytes = (double)n_boxes * box_bytes;
        printf("%-28s %7.1f B/clk/SM  %6.2f clk/row  %7.0f clk/box\n", v.name, bytes / avg, avg / ((double)n_boxes * v.brows), avg / n_boxes);
    }
    return 0;
}
