// stage_kernels.cu - device-side FITS data-unit staging.
//
// Observation.save_observation / save_mapped_observation (observation.py:1185-1303, :1315-1474)
// append one `fits.ImageHDU(data=<float64 array>)` per backplane to an HDUList and let astropy
// serialise it: every data unit is the array as big-endian IEEE-754 doubles, zero-padded to a
// multiple of 2880 bytes (FITS standard 4.0, sections 3.3.2 and 5.3).  With the backplanes
// already in HBM, the reference's "26 arrays -> 26 device-to-host copies -> 26 byte swaps on
// one host core" becomes: one launch that byte-swaps every array straight into its slot of a
// device-resident image of the whole file, then ONE copy of that image to pinned host memory.
//
// HBM-bound (8 B read + 8 B written per element, nothing reused): 8-byte accesses, fully
// coalesced, four independent elements in flight per thread, grid = a multiple of the SM count.
#include "pm_kernels.h"

namespace pm {

namespace {

constexpr int kStageBlock = 256;
constexpr int kStageUnroll = 4;

struct StageUnits {
    const double *src[PM_FITS_MAX_UNITS];
    unsigned long long *dst[PM_FITS_MAX_UNITS];  // first 8-byte word of the data unit in the image
    int64_t n_elems[PM_FITS_MAX_UNITS];          // doubles in the array
    int64_t word_end[PM_FITS_MAX_UNITS];         // exclusive prefix end, in 8-byte words incl. padding
    int n_units;
};

__device__ __forceinline__ unsigned long long bswap64(unsigned long long v) {
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    return ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | (unsigned long long)__byte_perm(hi, 0, 0x0123);
}

__global__ void __launch_bounds__(kStageBlock) fits_stage_kernel(const __grid_constant__ StageUnits units) {
    const int64_t total = units.word_end[units.n_units - 1];
    const int64_t stride = (int64_t)gridDim.x * kStageBlock;
    int u = 0;
    for (int64_t base = (int64_t)blockIdx.x * kStageBlock + threadIdx.x; base < total; base += stride * kStageUnroll) {
        unsigned long long v[kStageUnroll];
        unsigned long long *dst[kStageUnroll];
#pragma unroll
        for (int k = 0; k < kStageUnroll; k++) {
            const int64_t w = base + k * stride;
            dst[k] = nullptr;
            v[k] = 0ull;
            if (w < total) {
                while (w >= units.word_end[u]) u++;  // words only ever increase for one thread
                const int64_t local = w - (u ? units.word_end[u - 1] : 0);
                dst[k] = units.dst[u] + local;
                if (local < units.n_elems[u])
                    v[k] = __ldcs(reinterpret_cast<const unsigned long long *>(units.src[u]) + local);
            }
        }
#pragma unroll
        for (int k = 0; k < kStageUnroll; k++)
            if (dst[k]) __stcs(dst[k], bswap64(v[k]));
    }
}

}  // namespace

cudaError_t launch_fits_stage(const double *const *src, const int64_t *n_elems, const int64_t *dst_offset, int n_units,
                              uint8_t *image, int sm_count, cudaStream_t st) {
    for (int first = 0; first < n_units; first += PM_FITS_MAX_UNITS) {
        StageUnits units;
        units.n_units = 0;
        int64_t words = 0;
        for (int u = first; u < n_units && units.n_units < PM_FITS_MAX_UNITS; u++) {
            if (n_elems[u] == 0) continue;
            const int k = units.n_units++;
            units.src[k] = src[u];
            units.dst[k] = reinterpret_cast<unsigned long long *>(image + dst_offset[u]);
            units.n_elems[k] = n_elems[u];
            words += (n_elems[u] * 8 + 2879) / 2880 * 360;
            units.word_end[k] = words;
        }
        if (units.n_units == 0) continue;
        const int64_t per_cta = (int64_t)kStageBlock * kStageUnroll;
        int64_t ctas = (words + per_cta - 1) / per_cta;
        const int64_t cap = (int64_t)sm_count * 16;  // 8 resident CTAs per SM, two waves
        if (ctas > cap) ctas = cap;
        fits_stage_kernel<<<(unsigned)ctas, kStageBlock, 0, st>>>(units);
        count_launches(1);
        cudaError_t err = cudaGetLastError();
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

}  // namespace pm
