// gather_kernels.cu - cube -> map resampling (nearest / linear / not-a-knot cubic)
// and the spline preparation (NaN repair + B-spline coefficient solve).
//
// Replaces, for every wavelength plane at once, BodyXY._do_nearest_interpolation
// (planetmapper/body_xy.py:1633-1649), BodyXY._do_spline_interpolation
// (:1651-1702, scipy RectBivariateSpline s=0 -> FITPACK), _should_propagate_nan_to_map
// (:1855-1866) and _replace_nans_with_interpolated_values (:1871-1904), which
// Observation._get_mapped_data (planetmapper/observation.py:876-905) runs plane by
// plane in Python.
//
// Gather layout: thread = map cell (consecutive threads -> consecutive cells, so each
// plane's store is one coalesced 256 B warp transaction, written with a streaming
// hint because the output is never re-read); the per-cell interpolation weights are
// computed ONCE and reused for every wavelength plane of the CTA's plane group.
// The kernel is HBM-write bound: 8 B per output voxel.
#include "pm_device.cuh"
#include "pm_kernels.h"

#include <map>
#include <mutex>
#include <vector>

namespace pm {

constexpr int kGatherBlock = 256;
constexpr uint8_t kPlaneAllNan = 1;  // plane_skip bit0: np.all(np.isnan(img)) -> all NaN out
constexpr uint8_t kPlaneHasNan = 2;  // plane_skip bit1: some NaN pixel -> consult nanmask

// not-a-knot cubic knot vector for n samples at 0..n-1:
// t = [0,0,0,0, 2,3,...,n-3, n-1,n-1,n-1,n-1]   (FITPACK, s = 0)
__host__ __device__ __forceinline__ double nak_knot(int i, int n) {
    return i <= 3 ? 0.0 : (i >= n ? (double)(n - 1) : (double)(i - 2));
}

// FITPACK fpbspl: the 4 non-zero cubic B-splines at x; returns the first coefficient
// index.  x is clamped to [0, n-1] like bispev does.
__host__ __device__ __forceinline__ int bspline3_weights(double x, int n, double h[4]) {
    double xe = fmin(fmax(x, 0.0), (double)(n - 1));
    int j = (int)floor(xe);
    int l = j + 2;
    l = l < 3 ? 3 : (l > n - 1 ? n - 1 : l);
    double hh[3];
    h[0] = 1.0;
    h[1] = h[2] = h[3] = 0.0;
#pragma unroll
    for (int jj = 1; jj <= 3; jj++) {
#pragma unroll
        for (int i = 0; i < 3; i++)
            if (i < jj) hh[i] = h[i];
        h[0] = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            if (i < jj) {
                int li = l + 1 + i, lj = li - jj;
                double tli = nak_knot(li, n), tlj = nak_knot(lj, n);
                double f = hh[i] / (tli - tlj);
                h[i] = h[i] + f * (tli - xe);
                h[i + 1] = f * (xe - tlj);
            }
        }
    }
    return l - 3;
}

// ---------------------------------------------------------------------------------
// nearest neighbour: reads the raw cube [n_planes][ny][nx]
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGatherBlock)
    gather_nearest_kernel(const double *__restrict__ cube, int ny, int nx, int plane_begin, int plane_count,
                          const double *__restrict__ xmap, const double *__restrict__ ymap, int64_t n_cells,
                          double *__restrict__ out, int planes_per_group) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const int l0 = blockIdx.y * planes_per_group;
    const int l1 = min(l0 + planes_per_group, plane_count);
    const int64_t plane_px = (int64_t)ny * nx;
    const double x = __ldg(xmap + cell), y = __ldg(ymap + cell);
    const double nan = NAN;
    bool valid = !isnan(x);  // y is never NaN when x is not (body_xy.py:1646, :1697)
    int64_t off = 0;
    if (valid) {
        // np.round is round-half-to-even == rint (body_xy.py:1642-1643)
        long xi = (long)rint(x), yi = isnan(y) ? -999 : (long)rint(y);
        if (xi < 0) xi += nx;  // numpy negative indices wrap
        if (yi < 0) yi += ny;
        valid = xi >= 0 && xi < nx && yi >= 0 && yi < ny;
        off = (int64_t)yi * nx + xi;
    }
    const double *src = cube + (int64_t)(plane_begin + l0) * plane_px + off;
    double *dst = out + (int64_t)l0 * n_cells + cell;
    int l = l0;
    for (; l + 4 <= l1; l += 4) {
        double v0 = nan, v1 = nan, v2 = nan, v3 = nan;
        if (valid) {
            v0 = __ldg(src);
            v1 = __ldg(src + plane_px);
            v2 = __ldg(src + 2 * plane_px);
            v3 = __ldg(src + 3 * plane_px);
        }
        __stcs(dst, v0);
        __stcs(dst + n_cells, v1);
        __stcs(dst + 2 * n_cells, v2);
        __stcs(dst + 3 * n_cells, v3);
        src += 4 * plane_px;
        dst += 4 * n_cells;
    }
    for (; l < l1; l++) {
        __stcs(dst, valid ? __ldg(src) : nan);
        src += plane_px;
        dst += n_cells;
    }
}

// ---------------------------------------------------------------------------------
// spline modes: read the prepared operand of pm_spline_prepare
//   coefq   [ceil(n_planes / 4)][ny][nx][4]   four wavelength planes of one pixel are
//           32 contiguous bytes, so one 256-bit load feeds four FMAs and the K x K
//           footprint of a cell is K runs of K x 32 contiguous bytes;
//   nanbits [ny * nx][n_words]                bit (l % 32) of word l / 32 = pixel was NaN
//           in plane l: the propagate_nan test costs four word loads per 32 planes;
//   plane_bits [2][n_words]                   row 0: all-NaN planes, row 1: words with any
//           NaN pixel (nanbits needs consulting).
// The K x K weight products are formed once per cell; per output voxel the kernel then
// issues K*K/4 loads, K*K FMAs and one streaming store.
// ---------------------------------------------------------------------------------
struct Quad {
    double v[4];
};
__device__ __forceinline__ Quad ldg_quad(const double *p) {
    Quad q;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(q.v[0]), "=d"(q.v[1]), "=d"(q.v[2]), "=d"(q.v[3])
                 : "l"(p));
    return q;
}

template <int K>
__global__ void __launch_bounds__(kGatherBlock)
    gather_spline_kernel(const double *__restrict__ coefq, const uint32_t *__restrict__ nanbits,
                         const uint32_t *__restrict__ plane_bits, int n_words, int ny, int nx, int plane_begin,
                         int plane_count, const double *__restrict__ xmap, const double *__restrict__ ymap,
                         int64_t n_cells, uint32_t flags, double *__restrict__ out, int planes_per_group) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= n_cells) return;
    const int l0 = blockIdx.y * planes_per_group;  // relative to plane_begin, multiple of 4
    const int l1 = min(l0 + planes_per_group, plane_count);
    const double x = __ldg(xmap + cell), y = __ldg(ymap + cell);
    const double nan = NAN;
    bool valid = !isnan(x);  // y is never NaN when x is not (body_xy.py:1646, :1697)
    const bool propagate = (flags & PM_FLAG_PROPAGATE_NAN) != 0;
    int64_t nb00 = 0, nb01 = 0, nb10 = 0, nb11 = 0;
    if (valid && propagate) {
        // BodyXY._should_propagate_nan_to_map (body_xy.py:1855-1866)
        if (x < 0.0 || y < 0.0 || x > nx - 1 || y > ny - 1) valid = false;
        const int x0 = max((int)floor(x), 0), x1 = min((int)ceil(x), nx - 1);
        const int y0 = max((int)floor(y), 0), y1 = min((int)ceil(y), ny - 1);
        nb00 = ((int64_t)y0 * nx + x0) * n_words;
        nb01 = ((int64_t)y0 * nx + x1) * n_words;
        nb10 = ((int64_t)y1 * nx + x0) * n_words;
        nb11 = ((int64_t)y1 * nx + x1) * n_words;
    }
    double w[K * K];
    int ix = 0, iy = 0;
    if (valid) {
        double wx[K], wy[K];
        if (K == 4) {
            ix = bspline3_weights(x, nx, wx);
            iy = bspline3_weights(y, ny, wy);
        } else {
            const double xe = fmin(fmax(x, 0.0), (double)(nx - 1));
            const double ye = fmin(fmax(y, 0.0), (double)(ny - 1));
            ix = min((int)floor(xe), nx - 2);
            iy = min((int)floor(ye), ny - 2);
            const double fx = xe - ix, fy = ye - iy;
            wx[0] = 1.0 - fx;
            wx[1] = fx;
            wy[0] = 1.0 - fy;
            wy[1] = fy;
        }
#pragma unroll
        for (int a = 0; a < K; a++)
#pragma unroll
            for (int b = 0; b < K; b++) w[a * K + b] = wy[a] * wx[b];
    }
    const int64_t quad_stride = (int64_t)ny * nx * 4;  // doubles per plane quad
    const int64_t row_stride = (int64_t)nx * 4;
    const double *src = coefq + (int64_t)((plane_begin + l0) >> 2) * quad_stride + ((int64_t)iy * nx + ix) * 4;
    double *dst = out + (int64_t)l0 * n_cells + cell;
    uint32_t bad = 0;
    int cur_word = -1;
    for (int l = l0; l < l1; l += 4) {
        const int gl = plane_begin + l;  // global plane index of this quad (multiple of 4)
        const int word = gl >> 5;
        if (word != cur_word) {  // uniform: once per 32 planes
            cur_word = word;
            bad = __ldg(plane_bits + word);
            if (valid && propagate && __ldg(plane_bits + n_words + word))
                bad |= __ldg(nanbits + nb00 + word) | __ldg(nanbits + nb01 + word) | __ldg(nanbits + nb10 + word) |
                       __ldg(nanbits + nb11 + word);
        }
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        if (valid) {
            // FITPACK fpbisp order: rows (y) outer, columns inner
#pragma unroll
            for (int a = 0; a < K; a++) {
#pragma unroll
                for (int b = 0; b < K; b++) {
                    const Quad c = ldg_quad(src + a * row_stride + b * 4);
                    const double ww = w[a * K + b];
                    a0 = fma(c.v[0], ww, a0);
                    a1 = fma(c.v[1], ww, a1);
                    a2 = fma(c.v[2], ww, a2);
                    a3 = fma(c.v[3], ww, a3);
                }
            }
        }
        const uint32_t nib = valid ? ((bad >> (gl & 31)) & 0xFu) : 0xFu;
        const int left = l1 - l;
        __stcs(dst, (nib & 1u) ? nan : a0);
        if (left > 1) __stcs(dst + n_cells, (nib & 2u) ? nan : a1);
        if (left > 2) __stcs(dst + 2 * n_cells, (nib & 4u) ? nan : a2);
        if (left > 3) __stcs(dst + 3 * n_cells, (nib & 8u) ? nan : a3);
        src += quad_stride;
        dst += 4 * n_cells;
    }
}

cudaError_t launch_gather(const double *src, const uint32_t *nanbits, const uint32_t *plane_bits, int n_planes,
                          int ny, int nx, int plane_begin, int plane_count, const double *xmap, const double *ymap,
                          int64_t n_cells, int mode, uint32_t flags, double *out, int sm_count, cudaStream_t st) {
    (void)sm_count;
    if (n_cells == 0 || plane_count == 0) return cudaSuccess;
    int ppg = 128;  // planes per CTA: amortises the per-cell weights, bounds CTA run time
    if (plane_count < ppg) ppg = (plane_count + 3) / 4 * 4;
    dim3 grid((unsigned)((n_cells + kGatherBlock - 1) / kGatherBlock), (unsigned)((plane_count + ppg - 1) / ppg));
    const int n_words = (n_planes + 31) / 32;
    switch (mode) {
        case PM_INTERP_NEAREST:
            gather_nearest_kernel<<<grid, kGatherBlock, 0, st>>>(src, ny, nx, plane_begin, plane_count, xmap, ymap,
                                                                 n_cells, out, ppg);
            break;
        case PM_INTERP_LINEAR:
            gather_spline_kernel<2><<<grid, kGatherBlock, 0, st>>>(src, nanbits, plane_bits, n_words, ny, nx,
                                                                   plane_begin, plane_count, xmap, ymap, n_cells,
                                                                   flags, out, ppg);
            break;
        case PM_INTERP_CUBIC:
            gather_spline_kernel<4><<<grid, kGatherBlock, 0, st>>>(src, nanbits, plane_bits, n_words, ny, nx,
                                                                   plane_begin, plane_count, xmap, ymap, n_cells,
                                                                   flags, out, ppg);
            break;
        default:
            return cudaErrorInvalidValue;
    }
    count_launches(1);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------
// spline preparation
// ---------------------------------------------------------------------------------
struct PlaneStats {
    long long n_nan;
    long long n_bad;
};

// pass 1: NaN mask, bad-pixel counts, plane flags, copy into the coefficient buffer
__global__ void __launch_bounds__(256) classify_kernel(const double *__restrict__ cube, int64_t plane_px,
                                                       double *__restrict__ coef,
                                                       uint8_t *__restrict__ plane_skip,
                                                       PlaneStats *__restrict__ stats) {
    const int l = blockIdx.x;
    const double *src = cube + (int64_t)l * plane_px;
    double *dst = coef + (int64_t)l * plane_px;
    long long n_nan = 0, n_bad = 0;
    for (int64_t i = threadIdx.x; i < plane_px; i += blockDim.x) {
        double v = src[i];
        bool isn = isnan(v);
        n_nan += isn;
        n_bad += !isfinite(v);
        dst[i] = v;
    }
    __shared__ long long s_nan[256], s_bad[256];
    s_nan[threadIdx.x] = n_nan;
    s_bad[threadIdx.x] = n_bad;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_nan[threadIdx.x] += s_nan[threadIdx.x + o];
            s_bad[threadIdx.x] += s_bad[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        stats[l].n_nan = s_nan[0];
        stats[l].n_bad = s_bad[0];
        uint8_t flag = 0;
        if (s_nan[0] == plane_px) flag |= kPlaneAllNan;
        if (s_nan[0] > 0) flag |= kPlaneHasNan;
        plane_skip[l] = flag;
    }
}

__device__ __forceinline__ unsigned long long order_key(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_to_double(unsigned long long k) {
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// k-th smallest (0-based) finite value of a plane: MSB-first radix select, one CTA
__device__ double radix_select(const double *__restrict__ src, int64_t n, long long k) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix, s_mask;
    __shared__ long long s_k;
    if (threadIdx.x == 0) {
        s_prefix = 0;
        s_mask = 0;
        s_k = k;
    }
    __syncthreads();
    for (int pass = 7; pass >= 0; pass--) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix, mask = s_mask;
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            double v = src[i];
            if (isfinite(v)) {
                unsigned long long key = order_key(v);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> (8 * pass)) & 255ull], 1u);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long kk = s_k, cum = 0;
            int b = 0;
            for (; b < 256; b++) {
                if (cum + (long long)hist[b] > kk) break;
                cum += hist[b];
            }
            if (b > 255) b = 255;
            s_k = kk - cum;
            s_prefix = prefix | ((unsigned long long)b << (8 * pass));
            s_mask = mask | (0xffull << (8 * pass));
        }
        __syncthreads();
    }
    double r = key_to_double(s_prefix);
    __syncthreads();
    return r;
}

// pass 2: np.nanmedian of the finite pixels of each plane that has bad pixels
__global__ void __launch_bounds__(256) median_kernel(const double *__restrict__ cube, int64_t plane_px,
                                                     const PlaneStats *__restrict__ stats,
                                                     double *__restrict__ median) {
    const int l = blockIdx.x;
    const long long n_bad = stats[l].n_bad;
    if (n_bad == 0) return;  // nothing to repair (uniform for the CTA)
    const long long m = plane_px - n_bad;
    double med = 0.0;  // np.all(bad) -> 0.0 (body_xy.py:1890-1891)
    if (m > 0) {
        const double *src = cube + (int64_t)l * plane_px;
        if (m & 1) {
            med = radix_select(src, plane_px, m / 2);
        } else {
            double a = radix_select(src, plane_px, m / 2 - 1);
            double b = radix_select(src, plane_px, m / 2);
            med = (a + b) / 2.0;
        }
    }
    if (threadIdx.x == 0) median[l] = med;
}

// pass 3: replace bad pixels (body_xy.py:1893-1903)
__global__ void __launch_bounds__(256) repair_kernel(const double *__restrict__ cube, int n_planes, int ny,
                                                     int nx, const PlaneStats *__restrict__ stats,
                                                     const double *__restrict__ median,
                                                     double *__restrict__ coef) {
    const int64_t plane_px = (int64_t)ny * nx;
    const int64_t total = plane_px * n_planes;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)(idx / plane_px);
        if (stats[l].n_bad == 0) continue;
        const int64_t r = idx - (int64_t)l * plane_px;
        const int i = (int)(r / nx), j = (int)(r - (int64_t)i * nx);
        const double *img = cube + (int64_t)l * plane_px;
        if (isfinite(img[r])) continue;
        // uniform_filter(bad, size=3) on a bool array is True only when all nine
        // reflected neighbours are bad; reflect == clamp for size 3
        bool all_bad = true;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++) {
                int ii = min(max(i + di, 0), ny - 1), jj = min(max(j + dj, 0), nx - 1);
                all_bad = all_bad && !isfinite(img[(int64_t)ii * nx + jj]);
            }
        double v = median[l];
        if (!all_bad) {
            // np.nanmean over the window clipped at the image edge, inf treated as NaN
            double sum = 0.0;
            int cnt = 0;
            for (int ii = max(i - 1, 0); ii <= min(i + 1, ny - 1); ii++)
                for (int jj = max(j - 1, 0); jj <= min(j + 1, nx - 1); jj++) {
                    double w = img[(int64_t)ii * nx + jj];
                    if (isfinite(w)) {
                        sum += w;
                        cnt++;
                    }
                }
            v = sum / (double)cnt;
        }
        coef[idx] = v;
    }
}

// pass 4/5: separable not-a-knot B-spline coefficient solve with a banded (2,2) LU
// lu = [l1 | l2 | d | u1 | u2], each of length n
__device__ __forceinline__ void banded_solve(double *line, int n, int64_t stride, const double *__restrict__ lu) {
    const double *l1 = lu, *l2 = lu + n, *d = lu + 2 * n, *u1 = lu + 3 * n, *u2 = lu + 4 * n;
    double ym1 = 0.0, ym2 = 0.0;
    for (int i = 0; i < n; i++) {
        double yv = line[i * stride] - l1[i] * ym1 - l2[i] * ym2;
        line[i * stride] = yv;
        ym2 = ym1;
        ym1 = yv;
    }
    double xp1 = 0.0, xp2 = 0.0;
    for (int i = n - 1; i >= 0; i--) {
        double xv = (line[i * stride] - u1[i] * xp1 - u2[i] * xp2) / d[i];
        line[i * stride] = xv;
        xp2 = xp1;
        xp1 = xv;
    }
}
__global__ void __launch_bounds__(128) prefilter_rows_kernel(double *__restrict__ coef,
                                                             const uint8_t *__restrict__ plane_skip,
                                                             int n_planes, int ny, int nx,
                                                             const double *__restrict__ lu_x) {
    const int64_t lines = (int64_t)n_planes * ny;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < lines;
         t += (int64_t)gridDim.x * blockDim.x) {
        if (plane_skip[t / ny] & kPlaneAllNan) continue;
        banded_solve(coef + t * nx, nx, 1, lu_x);
    }
}
__global__ void __launch_bounds__(128) prefilter_cols_kernel(double *__restrict__ coef,
                                                             const uint8_t *__restrict__ plane_skip,
                                                             int n_planes, int ny, int nx,
                                                             const double *__restrict__ lu_y) {
    const int64_t lines = (int64_t)n_planes * nx;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < lines;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = t / nx;
        if (plane_skip[l] & kPlaneAllNan) continue;
        const int64_t col = t - l * nx;
        banded_solve(coef + l * (int64_t)ny * nx + col, ny, nx, lu_y);
    }
}

// pass 6: pack the repaired / prefiltered planes [l][y][x] into plane quads
// [l / 4][y][x][l % 4] (zero padding up to a multiple of four planes)
__global__ void __launch_bounds__(256) pack_quads_kernel(const double *__restrict__ planes, int n_planes,
                                                         int64_t plane_px, double *__restrict__ coefq) {
    const int64_t n_quads = (n_planes + 3) / 4;
    const int64_t total = n_quads * plane_px;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = t / plane_px, px = t - q * plane_px;
        double v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int64_t l = 4 * q + j;
            v[j] = l < n_planes ? planes[l * plane_px + px] : 0.0;
        }
        double4 o = make_double4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<double4 *>(coefq + t * 4) = o;
    }
}
// pass 7: NaN bit planes [pixel][word] of the ORIGINAL cube and the per-word plane bits
__global__ void __launch_bounds__(256) pack_nanbits_kernel(const double *__restrict__ cube, int n_planes,
                                                           int64_t plane_px, int n_words,
                                                           uint32_t *__restrict__ nanbits) {
    const int64_t total = plane_px * n_words;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t wd = t / plane_px, px = t - wd * plane_px;  // consecutive threads -> consecutive pixels
        uint32_t bits = 0;
        const int lbeg = (int)wd * 32, lend = min(lbeg + 32, n_planes);
        for (int l = lbeg; l < lend; l++) bits |= (isnan(cube[(int64_t)l * plane_px + px]) ? 1u : 0u) << (l - lbeg);
        nanbits[px * n_words + wd] = bits;
    }
}
__global__ void plane_bits_kernel(const uint8_t *__restrict__ plane_skip, int n_planes, int n_words,
                                  uint32_t *__restrict__ plane_bits) {
    const int wd = blockIdx.x * blockDim.x + threadIdx.x;
    if (wd >= n_words) return;
    uint32_t skip = 0, has = 0;
    for (int l = wd * 32; l < min(wd * 32 + 32, n_planes); l++) {
        if (plane_skip[l] & kPlaneAllNan) skip |= 1u << (l - wd * 32);
        if (plane_skip[l] & kPlaneHasNan) has = 1u;
    }
    plane_bits[wd] = skip;
    plane_bits[n_words + wd] = has;
}

// host: LU factors (no pivoting; the B-spline collocation matrix is totally positive)
static const std::vector<double> &nak_lu(int n) {
    static std::map<int, std::vector<double>> cache;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    // band storage: A[i][j - i + 2] for |j - i| <= 2
    std::vector<double> A((size_t)n * 5, 0.0);
    for (int i = 0; i < n; i++) {
        double h[4];
        int first = bspline3_weights((double)i, n, h);
        for (int m = 0; m < 4; m++) {
            int j = first + m;
            if (j < 0 || j >= n) continue;
            int off = j - i + 2;
            if (off >= 0 && off < 5) A[(size_t)i * 5 + off] = h[m];
        }
    }
    std::vector<double> lu((size_t)n * 5, 0.0);
    double *l1 = lu.data(), *l2 = l1 + n, *d = l2 + n, *u1 = d + n, *u2 = u1 + n;
    auto at = [&](int i, int j) -> double & { return A[(size_t)i * 5 + (j - i + 2)]; };
    for (int i = 0; i < n; i++) {
        for (int r = i + 1; r <= i + 2 && r < n; r++) {
            double m = at(r, i) / at(i, i);
            for (int j = i; j <= i + 2 && j < n; j++) {
                if (j - r >= -2 && j - r <= 2) at(r, j) -= m * at(i, j);
            }
            if (r == i + 1) l1[r] = m; else l2[r] = m;
        }
        d[i] = at(i, i);
        u1[i] = (i + 1 < n) ? at(i, i + 1) : 0.0;
        u2[i] = (i + 2 < n) ? at(i, i + 2) : 0.0;
    }
    return cache.emplace(n, std::move(lu)).first->second;
}

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

int64_t spline_coef_bytes(int n_planes, int ny, int nx) {
    return (int64_t)((n_planes + 3) / 4) * ny * nx * 4 * (int64_t)sizeof(double);
}
int64_t spline_nanbits_bytes(int n_planes, int ny, int nx) {
    return (int64_t)ny * nx * ((n_planes + 31) / 32) * (int64_t)sizeof(uint32_t);
}
int64_t spline_planebits_bytes(int n_planes) { return 2 * (int64_t)((n_planes + 31) / 32) * (int64_t)sizeof(uint32_t); }

int64_t spline_work_bytes(int n_planes, int ny, int nx, int degree) {
    int64_t b = align256((int64_t)n_planes * sizeof(PlaneStats)) + align256((int64_t)n_planes * sizeof(double)) +
                align256((int64_t)n_planes) + align256((int64_t)n_planes * ny * nx * (int64_t)sizeof(double));
    if (degree == 3) b += align256((int64_t)5 * nx * sizeof(double)) + align256((int64_t)5 * ny * sizeof(double));
    return b;
}

cudaError_t launch_spline_prepare(const double *cube, int n_planes, int ny, int nx, int degree, double *coefq,
                                  uint32_t *nanbits, uint32_t *plane_bits, void *work, int sm_count,
                                  cudaStream_t st) {
    if (n_planes == 0) return cudaSuccess;
    char *w = static_cast<char *>(work);
    PlaneStats *stats = reinterpret_cast<PlaneStats *>(w);
    w += align256((int64_t)n_planes * sizeof(PlaneStats));
    double *median = reinterpret_cast<double *>(w);
    w += align256((int64_t)n_planes * sizeof(double));
    uint8_t *plane_skip = reinterpret_cast<uint8_t *>(w);
    w += align256((int64_t)n_planes);
    double *coef = reinterpret_cast<double *>(w);  // natural layout [l][y][x] scratch
    w += align256((int64_t)n_planes * ny * nx * (int64_t)sizeof(double));
    const int64_t plane_px = (int64_t)ny * nx;
    const int n_words = (n_planes + 31) / 32;
    classify_kernel<<<n_planes, 256, 0, st>>>(cube, plane_px, coef, plane_skip, stats);
    median_kernel<<<n_planes, 256, 0, st>>>(cube, plane_px, stats, median);
    int64_t total = plane_px * n_planes;
    int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count * 16);
    repair_kernel<<<blocks, 256, 0, st>>>(cube, n_planes, ny, nx, stats, median, coef);
    count_launches(3);
    if (degree == 3) {
        double *lu_x = reinterpret_cast<double *>(w);
        w += align256((int64_t)5 * nx * sizeof(double));
        double *lu_y = reinterpret_cast<double *>(w);
        const std::vector<double> &hx = nak_lu(nx);
        const std::vector<double> &hy = nak_lu(ny);
        cudaError_t e = cudaMemcpyAsync(lu_x, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        e = cudaMemcpyAsync(lu_y, hy.data(), hy.size() * sizeof(double), cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
        int64_t rows = (int64_t)n_planes * ny, cols = (int64_t)n_planes * nx;
        int rb = (int)std::min<int64_t>((rows + 127) / 128, (int64_t)sm_count * 16);
        int cb = (int)std::min<int64_t>((cols + 127) / 128, (int64_t)sm_count * 16);
        prefilter_rows_kernel<<<rb, 128, 0, st>>>(coef, plane_skip, n_planes, ny, nx, lu_x);
        prefilter_cols_kernel<<<cb, 128, 0, st>>>(coef, plane_skip, n_planes, ny, nx, lu_y);
        count_launches(2);
    }
    const int64_t quads_total = (int64_t)((n_planes + 3) / 4) * plane_px;
    pack_quads_kernel<<<(int)std::min<int64_t>((quads_total + 255) / 256, (int64_t)sm_count * 32), 256, 0, st>>>(
        coef, n_planes, plane_px, coefq);
    const int64_t words_total = plane_px * n_words;
    pack_nanbits_kernel<<<(int)std::min<int64_t>((words_total + 255) / 256, (int64_t)sm_count * 32), 256, 0, st>>>(
        cube, n_planes, plane_px, n_words, nanbits);
    plane_bits_kernel<<<(n_words + 127) / 128, 128, 0, st>>>(plane_skip, n_planes, n_words, plane_bits);
    count_launches(3);
    return cudaGetLastError();
}

}  // namespace pm
