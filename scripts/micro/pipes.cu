// Micro-benchmark: issue rates of F2F.F64.F32 (XU), DFMA (FP64) and their mix on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, const float* in, int iters)
{
    float f[8];
    for (int i = 0; i < 8; ++i) f[i] = in[threadIdx.x + 32 * i];
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double w = in[0] + 1.000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {  // cvt + fma (the gather's inner pair)
                acc[i] = fma((double)f[i], w, acc[i]);
                f[i] = __int_as_float(__float_as_int(f[i]) + 1);  // keep the cvt inside the loop
            } else if (MODE == 1) {  // fma only
                acc[i] = fma(acc[i], w, 1.0);
                f[i] = __int_as_float(__float_as_int(f[i]) + 1);
            } else if (MODE == 2) {  // cvt only (sum in FP32 domain is not possible: use int xor on the result)
                double d = (double)f[i];
                acc[i] = __longlong_as_double(__double_as_longlong(acc[i]) ^ __double_as_longlong(d));
                f[i] = __int_as_float(__float_as_int(f[i]) + 1);
            } else {  // 2 fma per cvt
                acc[i] = fma((double)f[i], w, acc[i]);
                acc[i] = fma(acc[i], w, 0.5);
                f[i] = __int_as_float(__float_as_int(f[i]) + 1);
            }
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += acc[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
}

template <int MODE>
void run(const char* name, int warps_per_sm, int sms)
{
    float *in, *out;
    cudaMalloc(&in, 4096 * 4);
    cudaMemset(in, 0, 4096 * 4);
    cudaMalloc(&out, (size_t)sms * warps_per_sm * 32 * 4);
    const int iters = 20000;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, in, 100);
    cudaEventRecord(a);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, in, iters);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk * 1e3;
    double per_sm_units = (double)warps_per_sm * iters * 8;  // warp-level "units" per SM
    printf("%-22s warps/SM %2d: %.2f cycles per warp-unit per SM (%.3f ms)\n", name, warps_per_sm,
           cycles / per_sm_units, ms);
    cudaFree(in);
    cudaFree(out);
}

int main()
{
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int w : {4, 8, 16, 32}) {
        run<0>("cvt+fma", w, sms);
        run<1>("fma", w, sms);
        run<2>("cvt(+2 lop)", w, sms);
        run<3>("cvt+2fma", w, sms);
    }
    return 0;
}
