// fields.cu -- field store: gradient pre-pass + packing into the push kernel's records.
//
// Replaces calc_fields_gradients (mhd_data_parallel.f90:504-605) for uniform Cartesian
// grids and the consumer side of read_field_data_parallel (mhd_data_parallel.f90:224).
// The arithmetic is the reference's: FP32 difference, times 0.5/dx in FP64, rounded back
// to FP32 on store; one-sided 3-point formula at the two ends of each array axis.  Only the
// slots the configured pusher reads are materialised (gpat_internal.cuh, Rec<L>).
#include "gpat_internal.cuh"

namespace gpat {

struct SlotMap {
    int nrec;
    int slot[32];  // reference slot (1-based) for each packed position, 0 = padding
};

struct GridDims {
    int nxg, nyg, nzg;
    double idxh, idyh, idzh;  // 0.5/dx, 0.5/dy, 0.5/dz (mhd_data_parallel.f90:512-514)
};

// gradient of primary v along direction d (0,1,2) at storage cell (i,j,k); src has `nvar`
// floats per cell with the primaries first.
__device__ __forceinline__ float grad_one(const float* __restrict__ src, int nvar, const GridDims& g,
                                          int v, int d, int i, int j, int k)
{
    const long long sx = nvar, sy = (long long)nvar * g.nxg, sz = (long long)nvar * g.nxg * g.nyg;
    const float* c = src + (long long)i * sx + (long long)j * sy + (long long)k * sz + v;
    long long s;
    int pos, n;
    double idh;
    if (d == 0) { s = sx; pos = i; n = g.nxg; idh = g.idxh; }
    else if (d == 1) { s = sy; pos = j; n = g.nyg; idh = g.idyh; }
    else { s = sz; pos = k; n = g.nzg; idh = g.idzh; }
    if (n <= 1) return 0.0f;  // unresolved dimension: the reference leaves the zero fill
    float diff;
    if (pos == 0) {  // mhd_data_parallel.f90:537-539
        float a = __fmul_rn(-3.0f, c[0]);
        float b = __fmul_rn(4.0f, c[s]);
        diff = __fsub_rn(__fadd_rn(a, b), c[2 * s]);
    } else if (pos == n - 1) {  // mhd_data_parallel.f90:540-542
        float a = __fmul_rn(3.0f, c[0]);
        float b = __fmul_rn(4.0f, c[-s]);
        diff = __fadd_rn(__fsub_rn(a, b), c[-2 * s]);
    } else {  // mhd_data_parallel.f90:535-536
        diff = __fsub_rn(c[s], c[-s]);
    }
    return __double2float_rn(__dmul_rn((double)diff, idh));
}

// One thread per grid point: fills its record (one frame half) in the packed store.  A grid
// point owns 2*nrec floats made of 32-byte chunks [4 slots of half 0 | the same 4 slots of
// half 1] (gpat_internal.cuh "packed field record layouts"): both time frames of a slot quad
// arrive with one 256-bit load.
__global__ void pack_kernel(const float* __restrict__ src, int nvar, int with_grad, GridDims g,
                            SlotMap map, float* __restrict__ dst, long long stride, int half_off)
{
    const long long ncell = (long long)g.nxg * g.nyg * g.nzg;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
         c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % g.nxg);
        const int j = (int)((c / g.nxg) % g.nyg);
        const int k = (int)(c / ((long long)g.nxg * g.nyg));
        float* out = dst + c * stride + half_off;
        for (int q = 0; q < map.nrec; q += 4) {
            float v4[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int s = map.slot[q + e];
                float val = 0.0f;
                if (s >= 1 && s <= 8) val = src[c * nvar + (s - 1)];
                else if (s > 8) {
                    if (with_grad) val = src[c * nvar + (s - 1)];
                    else val = grad_one(src, nvar, g, (s - 9) / 3, (s - 9) % 3, i, j, k);
                }
                v4[e] = val;
            }
            *reinterpret_cast<float4*>(out + 2 * q) = make_float4(v4[0], v4[1], v4[2], v4[3]);
        }
    }
}

// Side plane of L2D (gpat_internal.cuh): one float4 per grid point, [s0, s1 of half 0 | s0, s1 of half 1];
// this frame half owns two of the four floats.  Same gradient arithmetic as the record plane.
__global__ void pack_side_kernel(const float* __restrict__ src, int nvar, int with_grad, GridDims g, int s0, int s1,
                                 float* __restrict__ side, int half)
{
    const long long ncell = (long long)g.nxg * g.nyg * g.nzg;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
         c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % g.nxg);
        const int j = (int)((c / g.nxg) % g.nyg);
        const int k = (int)(c / ((long long)g.nxg * g.nyg));
        float2 v;
        if (with_grad) { v.x = src[c * nvar + (s0 - 1)]; v.y = src[c * nvar + (s1 - 1)]; }
        else {
            v.x = grad_one(src, nvar, g, (s0 - 9) / 3, (s0 - 9) % 3, i, j, k);
            v.y = grad_one(src, nvar, g, (s1 - 9) / 3, (s1 - 9) % 3, i, j, k);
        }
        *reinterpret_cast<float2*>(side + c * 4 + 2 * half) = v;
    }
}

// debug: the full 32-slot reference layout from an 8-variable frame
__global__ void grad32_kernel(const float* __restrict__ src, GridDims g, float* __restrict__ out32)
{
    const long long ncell = (long long)g.nxg * g.nyg * g.nzg;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < ncell;
         c += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(c % g.nxg);
        const int j = (int)((c / g.nxg) % g.nyg);
        const int k = (int)(c / ((long long)g.nxg * g.nyg));
        for (int v = 0; v < 8; ++v) out32[c * 32 + v] = src[c * 8 + v];
        for (int s = 9; s <= 32; ++s)
            out32[c * 32 + s - 1] = grad_one(src, 8, g, (s - 9) / 3, (s - 9) % 3, i, j, k);
    }
}

void launch_pack(const float* src, int nvar, int with_grad, const DevParams& prm, int layout,
                 float* dst, int half, int sm_count, cudaStream_t st)
{
    // source and destination share the cell index: a 1-D frame fills row 0 of the two-row store
    GridDims g{prm.nxg, prm.nyg_src, prm.nzg, 0.5 / prm.dx, 0.5 / prm.dy, 0.5 / prm.dz};
    SlotMap map;
    map.nrec = nrec_of(layout);
    for (int k = 0; k < 32; ++k) map.slot[k] = (k < map.nrec) ? slot_of(layout, k) : 0;
    const long long stride = 2LL * map.nrec;
    const int half_off = (prm.time_interp ? half : 0) * 4;
    pack_kernel<<<sm_count * 8, 256, 0, st>>>(src, nvar, with_grad, g, map, dst, stride, half_off);
    if (layout == L3D) {  // side plane in chunk format: the same kernel over the side slots, 16 floats per grid point
        SlotMap sm;
        sm.nrec = 8;
        for (int k = 0; k < 32; ++k) sm.slot[k] = (k < 8) ? side_slot_of(layout, k) : 0;
        const long long nstore = (long long)prm.nxg * prm.nyg * prm.nzg;
        pack_kernel<<<sm_count * 8, 256, 0, st>>>(src, nvar, with_grad, g, sm, dst + nstore * stride, 16, half_off);
    } else if (side_floats_of(layout)) {
        const long long nstore = (long long)prm.nxg * prm.nyg * prm.nzg;  // the side plane starts behind the records
        pack_side_kernel<<<sm_count * 8, 256, 0, st>>>(src, nvar, with_grad, g, side_slot_of(layout, 0),
                                                       side_slot_of(layout, 1), dst + nstore * stride,
                                                       prm.time_interp ? half : 0);
    }
}

// Turbulence maps: consumer side of read_magnetic_fluctuation / read_correlation_length
// (mhd_data_parallel.f90:306-497) + calc_grad_sigma2_slab/_2d, calc_grad_lc_slab/_2d (:771-1604).
// src holds the slab array then the 2-D array (each one float per ghosted grid point); chunk
// 2*which + t of every grid point receives [value, d/dx, d/dy, d/dz] of array t for this frame half.
__global__ void pack_aux_kernel(const float* __restrict__ src, GridDims g, int which, float* __restrict__ aux,
                                int half_off)
{
    const long long ncell = (long long)g.nxg * g.nyg * g.nzg;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < 2 * ncell;
         c += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(c / ncell);
        const long long cc = c % ncell;
        const int i = (int)(cc % g.nxg);
        const int j = (int)((cc / g.nxg) % g.nyg);
        const int k = (int)(cc / ((long long)g.nxg * g.nyg));
        const float* a = src + t * ncell;
        float4 v;
        v.x = a[cc];
        // an unresolved axis keeps the 1.0 the reference initialises these arrays with (:125-126)
        v.y = (g.nxg > 1) ? grad_one(a, 1, g, 0, 0, i, j, k) : 1.0f;
        v.z = (g.nyg > 1) ? grad_one(a, 1, g, 0, 1, i, j, k) : 1.0f;
        v.w = (g.nzg > 1) ? grad_one(a, 1, g, 0, 2, i, j, k) : 1.0f;
        *reinterpret_cast<float4*>(aux + cc * 32 + (2 * which + t) * 8 + half_off) = v;
    }
}

__global__ void fill_kernel(float* __restrict__ p, long long n, float v)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}
void launch_fill(float* p, long long n, float v, int sm_count, cudaStream_t st)
{
    fill_kernel<<<sm_count * 8, 256, 0, st>>>(p, n, v);
}

void launch_pack_aux(const float* src2, const DevParams& prm, int which, float* aux, int half, int sm_count,
                     cudaStream_t st)
{
    GridDims g{prm.nxg, prm.nyg_src, prm.nzg, 0.5 / prm.dx, 0.5 / prm.dy, 0.5 / prm.dz};
    pack_aux_kernel<<<sm_count * 8, 256, 0, st>>>(src2, g, which, aux, (prm.time_interp ? half : 0) * 4);
}

void launch_grad32(const float* src8, const DevParams& prm, float* out32, int sm_count,
                   cudaStream_t st)
{
    GridDims g{prm.nxg, prm.nyg_src, prm.nzg, 0.5 / prm.dx, 0.5 / prm.dy, 0.5 / prm.dz};
    grad32_kernel<<<sm_count * 8, 256, 0, st>>>(src8, g, out32);
}

}  // namespace gpat
