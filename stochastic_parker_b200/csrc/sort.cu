// sort.cu -- spatial ordering of the particle arrays before a push (production build).
//
// The reference keeps particles in injection / swap-with-tail order (particle_module.f90:454-530,
// 5365-5403); nothing in its arithmetic depends on that order -- every particle owns its random
// stream here -- so the library is free to choose the order that is best for the memory system.
// Particles are sorted by the cell they sit in (x fastest) at the start of gpat_particle_mover: the
// lanes of a warp, which take consecutive particles from the work queue, then gather from
// neighbouring grid points and share their 128-byte lines.  Measured (profiles/README.md r01j):
// C5 (3-D, store does not fit the L2) 2.8e9 -> 5.7e9 steps/s, C1 / C2 +4 %, C4 +10 %, sort included;
// C3, whose 68 MB store is L2-resident as a whole, loses 4.5 % (particles with similar step counts end
// up in the same warps), so the default sorts only when the store is larger than the L2.
// Keys: one kernel; order: cub::DeviceRadixSort (CCCL, shipped with the CUDA toolkit -- library
// plumbing, not the hot path); permutation: one gather kernel over the 17 SoA arrays into the
// second particle buffer, after which the two buffers swap roles.
// The reference-order build never sorts: its tests check the reference's own particle order.
#include <cub/device/device_radix_sort.cuh>

#include "gpat_internal.cuh"

namespace gpat {

__global__ void cell_key_kernel(const __grid_constant__ DevParams prm, const double* __restrict__ x,
                                const double* __restrict__ y, const double* __restrict__ z, long long n,
                                unsigned* __restrict__ key, unsigned* __restrict__ idx)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = (int)floor((x[i] - prm.xmin) * prm.idx);
    int cy = (int)floor((y[i] - prm.ymin) * prm.idy);
    int cz = (prm.ndim == 3) ? (int)floor((z[i] - prm.zmin) * prm.idz) : 0;
    cx = min(max(cx, 0), prm.nx - 1);
    cy = min(max(cy, 0), prm.ny - 1);
    cz = min(max(cz, 0), prm.nz - 1);
    key[i] = (unsigned)(((long long)cz * prm.ny + cy) * prm.nx + cx);
    idx[i] = (unsigned)i;
}

__global__ void permute_kernel(PtlSoA D, PtlSoA S, const unsigned* __restrict__ idx, long long n)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long s = idx[i];
    D.x[i] = S.x[s]; D.y[i] = S.y[s]; D.z[i] = S.z[s]; D.p[i] = S.p[s]; D.v[i] = S.v[s];
    D.mu[i] = S.mu[s]; D.weight[i] = S.weight[s]; D.t[i] = S.t[s]; D.dt[i] = S.dt[s];
    D.rng[i] = S.rng[s]; D.origin[i] = S.origin[s]; D.nsteps_tracked[i] = S.nsteps_tracked[s];
    D.nsteps_pushed[i] = S.nsteps_pushed[s]; D.tag_injected[i] = S.tag_injected[s];
    D.tag_splitted[i] = S.tag_splitted[s]; D.split_times[i] = S.split_times[s];
    D.count_flag[i] = S.count_flag[s];
}

size_t sort_scratch_bytes(long long n)
{
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned*)nullptr, (unsigned*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)n);
    return tmp;
}

// keys/idx: 2 x n unsigned each (in, out); tmp: sort_scratch_bytes(n).  Leaves the sorted
// particles in D (the caller swaps the buffers).
cudaError_t launch_cell_sort(const DevParams& prm, const PtlSoA& S, const PtlSoA& D, long long n, unsigned* keys,
                             unsigned* idx, void* tmp, size_t tmp_bytes, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cell_key_kernel<<<grid, 256, 0, st>>>(prm, S.x, S.y, S.z, n, keys, idx);
    const long long ncell = (long long)prm.nx * prm.ny * prm.nz;
    int bits = 1;
    while ((1LL << bits) < ncell && bits < 32) ++bits;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys + n, idx, idx + n, (int)n, 0, bits, st);
    if (e != cudaSuccess) return e;
    permute_kernel<<<grid, 256, 0, st>>>(D, S, idx + n, n);
    return cudaGetLastError();
}

}  // namespace gpat
