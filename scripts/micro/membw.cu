// libmembw.so -- measured read bandwidth of L2 and of HBM with the push kernel's own load instruction
// (ld.global.nc.v8.f32 = LDG.E.256), for bench.py's roofline ("the binding ceiling of C1 is not HBM":
// VERDICT r01).  Measurement tooling, not part of the C ABI of the product.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o libmembw.so membw.cu
//   int membw_read_gbs(size_t bytes, int passes, int mode, double* gbs)
//     mode 0: every CTA streams the whole buffer `passes` times (stride = grid), coalesced 32 B per lane;
//     mode 1: each lane GROUP of 4 reads one 128-byte line at a pseudo-random line index (the push
//             kernel's gather pattern: 8 distinct lines per warp-wide load).
//   A buffer well below the 126 MB L2 measures L2, a buffer of several GB measures HBM.
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ldg256(const float* p, float4& lo, float4& hi)
{
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w)
                 : "l"(p));
}

template <int MODE>
__global__ void __launch_bounds__(256) read_kernel(const float* __restrict__ buf, size_t nchunks, int passes, float* sink)
{
    // a chunk = 32 bytes = 8 floats
    float acc = 0.f;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nthreads = (size_t)gridDim.x * blockDim.x;
    if (MODE == 0) {
        for (int p = 0; p < passes; ++p) {
            size_t c = tid;
            // 4 independent loads in flight per lane
            for (; c + 3 * nthreads < nchunks; c += 4 * nthreads) {
                float4 a0, b0, a1, b1, a2, b2, a3, b3;
                ldg256(buf + 8 * c, a0, b0);
                ldg256(buf + 8 * (c + nthreads), a1, b1);
                ldg256(buf + 8 * (c + 2 * nthreads), a2, b2);
                ldg256(buf + 8 * (c + 3 * nthreads), a3, b3);
                acc += a0.x + b0.w + a1.y + b1.z + a2.z + b2.y + a3.w + b3.x;
            }
            for (; c < nchunks; c += nthreads) {
                float4 a0, b0;
                ldg256(buf + 8 * c, a0, b0);
                acc += a0.x + b0.w;
            }
        }
    } else {
        // random 128-byte lines: the line count is rounded down to a power of two so that the index is a mask (a 64-bit
        // modulo would make the loop compute-bound), eight independent loads in flight per lane
        size_t nlines = 1;
        while (nlines * 2 <= nchunks / 4) nlines *= 2;
        const size_t mask = nlines - 1;
        const unsigned q = threadIdx.x & 3u;
        uint64_t s = (tid >> 2) * 0x9E3779B97F4A7C15ull + 12345u;
        const size_t per_thread = (nchunks * (size_t)passes) / nthreads;
        for (size_t i = 0; i + 7 < per_thread; i += 8) {
            float4 a[8], b[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s = s * 6364136223846793005ull + 1442695040888963407ull;
                const size_t line = (size_t)(s >> 24) & mask;
                ldg256(buf + 32 * line + 8 * q, a[j], b[j]);
            }
            acc += a[0].x + b[1].y + a[2].z + b[3].w + a[4].x + b[5].y + a[6].z + b[7].w;
        }
    }
    if (acc == 123.456f) sink[0] = acc;
}

extern "C" int membw_read_gbs(size_t bytes, int passes, int mode, double* gbs)
{
    float *buf = nullptr, *sink = nullptr;
    bytes = (bytes / 4096) * 4096;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) return 1;
    cudaMalloc(&sink, 64);
    cudaMemset(buf, 0, bytes);
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = sms * 8;
    const size_t nchunks = bytes / 32;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms the L2
        cudaEventRecord(e0);
        if (mode == 0) read_kernel<0><<<grid, 256>>>(buf, nchunks, passes, sink);
        else read_kernel<1><<<grid, 256>>>(buf, nchunks, passes, sink);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 2;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    size_t moved;
    if (mode == 0) moved = bytes * (size_t)passes;
    else {
        const size_t nthreads = (size_t)grid * 256;
        moved = ((nchunks * (size_t)passes) / nthreads / 8 * 8) * nthreads * 32;
    }
    *gbs = (double)moved / (best * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    cudaFree(sink);
    return cudaGetLastError() == cudaSuccess ? 0 : 3;
}
