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
// 3-D keys are TILED (round 2): tile index first, then the cell inside the tile (x fastest).  The resident
// lanes hold a contiguous range of ~8e4-1.5e5 particles of the sorted array; in plain x-y-z order that range
// spans two whole planes of the grid (512^2 x 192 B x 2 = 100 MB at C5's size: the L2 misses and the gathers run
// at the random-line HBM rate, 4.8 TB/s measured); in 256 x 64 x 32 tiles (524 288 cells) it sits inside one or two
// tiles whose x-rows are still 48 KB of contiguous records.  Measured on C5 at full size (512^3, 1.25e8 particles,
// profiles/r02h_c5_tiles_512.log): untiled 5.64e9 steps/s, 64^3 5.66e9 (rows too short), 512x32x8 6.86e9,
// 128x128x64 6.91e9, 256x64x32 6.98e9.  GPAT_SORT_TILE="tx,ty,tz" overrides (0 = untiled).
// The reference-order build never sorts: its tests check the reference's own particle order.
#include <cstdio>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "gpat_internal.cuh"

namespace gpat {

struct SortTile { int tx, ty, tz, ntx, nty; };  // tx = 0: untiled

__global__ void cell_key_kernel(const __grid_constant__ DevParams prm, const double* __restrict__ x,
                                const double* __restrict__ y, const double* __restrict__ z, long long n,
                                unsigned* __restrict__ key, unsigned* __restrict__ idx, SortTile tl)
{
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = (int)floor((x[i] - prm.xmin) * prm.idx);
    int cy = (int)floor((y[i] - prm.ymin) * prm.idy);
    int cz = (prm.ndim == 3) ? (int)floor((z[i] - prm.zmin) * prm.idz) : 0;
    cx = min(max(cx, 0), prm.nx - 1);
    cy = min(max(cy, 0), prm.ny - 1);
    cz = min(max(cz, 0), prm.nz - 1);
    if (tl.tx > 0) {
        const int bx = cx / tl.tx, by = cy / tl.ty, bz = cz / tl.tz;
        const long long tile = ((long long)bz * tl.nty + by) * tl.ntx + bx;
        const long long in = ((long long)(cz - bz * tl.tz) * tl.ty + (cy - by * tl.ty)) * tl.tx + (cx - bx * tl.tx);
        key[i] = (unsigned)(tile * ((long long)tl.tx * tl.ty * tl.tz) + in);
    } else {
        key[i] = (unsigned)(((long long)cz * prm.ny + cy) * prm.nx + cx);
    }
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
    SortTile tl{0, 1, 1, 1, 1};
    if (prm.ndim == 3) { tl.tx = 256; tl.ty = 64; tl.tz = 32; }  // profiles/r02h_c5_tiles_512.log
    if (const char* e = getenv("GPAT_SORT_TILE")) {
        int a = 0, b = 1, c = 1;
        if (sscanf(e, "%d,%d,%d", &a, &b, &c) >= 1) { tl.tx = a; tl.ty = b > 0 ? b : 1; tl.tz = c > 0 ? c : 1; }
    }
    long long ncell = (long long)prm.nx * prm.ny * prm.nz;
    if (tl.tx > 0) {
        tl.ntx = (prm.nx + tl.tx - 1) / tl.tx;
        tl.nty = (prm.ny + tl.ty - 1) / tl.ty;
        const long long ntz = (prm.nz + tl.tz - 1) / tl.tz;
        ncell = (long long)tl.ntx * tl.nty * ntz * tl.tx * tl.ty * tl.tz;
        if (ncell >= (1LL << 32)) { tl.tx = 0; ncell = (long long)prm.nx * prm.ny * prm.nz; }  // keys are 32-bit
    }
    cell_key_kernel<<<grid, 256, 0, st>>>(prm, S.x, S.y, S.z, n, keys, idx, tl);
    int bits = 1;
    while ((1LL << bits) < ncell && bits < 32) ++bits;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys + n, idx, idx + n, (int)n, 0, bits, st);
    if (e != cudaSuccess) return e;
    permute_kernel<<<grid, 256, 0, st>>>(D, S, idx + n, n);
    return cudaGetLastError();
}

}  // namespace gpat
