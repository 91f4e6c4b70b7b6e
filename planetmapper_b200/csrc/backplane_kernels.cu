// backplane_kernels.cu - fused per-pixel (image direction) and per-cell (map
// direction) backplane kernels, plus the vectorised point transforms.
//
// One thread per pixel / cell; consecutive threads take consecutive x (or cell index)
// so every plane store is a fully coalesced 256-byte warp transaction.  A block
// first stages its frame's PMFrame (736 B) and derived constants in shared memory.
// The image kernel is launched as grid (tiles, n_frames): frames of a time series
// are batched into a single launch.
#include "pm_device.cuh"
#include "pm_kernels.h"

namespace pm {

__host__ __device__ constexpr uint64_t bit(int k) { return 1ull << k; }
constexpr uint64_t kKmMask = bit(PM_KM_X) | bit(PM_KM_Y) | bit(PM_ANGULAR_X) | bit(PM_ANGULAR_Y);
constexpr uint64_t kLonLatMask = bit(PM_LON_GRAPHIC) | bit(PM_LAT_GRAPHIC) | bit(PM_LOCAL_SOLAR_TIME);
constexpr uint64_t kCentricMask = bit(PM_LON_CENTRIC) | bit(PM_LAT_CENTRIC);
constexpr uint64_t kIllumMask = bit(PM_PHASE) | bit(PM_INCIDENCE) | bit(PM_EMISSION) | bit(PM_AZIMUTH);
constexpr uint64_t kStateMask = bit(PM_DISTANCE) | bit(PM_RADIAL_VELOCITY) | bit(PM_DOPPLER);
constexpr uint64_t kLimbMask = bit(PM_LIMB_DISTANCE) | bit(PM_LIMB_LON_GRAPHIC) | bit(PM_LIMB_LAT_GRAPHIC);
constexpr uint64_t kRingMask = bit(PM_RING_RADIUS) | bit(PM_RING_LON_GRAPHIC) | bit(PM_RING_DISTANCE);
constexpr uint64_t kSurfMask = kLonLatMask | kCentricMask | kIllumMask | kStateMask | kRingMask;

// plane slot of id k inside the packed output = number of requested planes below k
__device__ __forceinline__ int slot(uint64_t mask, int k) { return __popcll(mask & (bit(k) - 1ull)); }

#define PM_STORE(k, val)                                                            \
    do {                                                                            \
        if (mask & bit(k)) __stcs(out + (int64_t)slot(mask, k) * plane_stride + idx, (val)); \
    } while (0)

// ---------------------------------------------------------------------------------
// Image direction: all default backplanes for every pixel of every frame.
// Replaces the loops listed at pm_backplanes_img in include/pm_b200.h.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) backplanes_img_kernel(const PMFrame *__restrict__ frames,
                                                                int nx, int ny, uint64_t mask,
                                                                double *__restrict__ out_all) {
    __shared__ FrameS fs;
    load_frame(fs, frames + blockIdx.y);
    const PMFrame &f = fs.f;
    const int64_t npx = (int64_t)nx * ny;
    const int64_t plane_stride = npx;
    double *out = out_all + (int64_t)blockIdx.y * __popcll(mask) * npx;
    const double nan = NAN;

    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < npx;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int yi = (int)(idx / nx), xi = (int)(idx - (int64_t)yi * nx);
        const double x = (double)xi, y = (double)yi;
        PM_STORE(PM_PIXEL_X, x);  // BodyXY.get_x_img / get_y_img (body_xy.py:3494-3531)
        PM_STORE(PM_PIXEL_Y, y);

        // BodyXY._get_radec_img (body_xy.py:3413-3418)
        V3 d = xy2obsvec_norm(f, x, y);
        V3 d2 = d;
        if (mask & (bit(PM_RA) | bit(PM_DEC) | kKmMask | kLimbMask | kRingMask)) {
            double ra, dec;
            recrad_angles(d, ra, dec);
            const double ra_deg = ra * kDpr, dec_deg = dec * kDpr;
            PM_STORE(PM_RA, ra_deg);
            PM_STORE(PM_DEC, dec_deg);
            // _get_obsvec_norm_img (body_xy.py:3263-3272): RA/Dec degrees -> unit vector
            if (mask & (kKmMask | kLimbMask | kRingMask)) d2 = radrec1(ra_deg * kRpd, dec_deg * kRpd);
        }
        if (mask & kKmMask) {  // _get_km_xy_img (body_xy.py:3547-3553), angular (:3611-3656)
            double kx, ky;
            obsvec2km(f, d2, kx, ky);
            PM_STORE(PM_KM_X, kx);
            PM_STORE(PM_KM_Y, ky);
            PM_STORE(PM_ANGULAR_X, kx / f.km_per_arcsec);
            PM_STORE(PM_ANGULAR_Y, ky / f.km_per_arcsec);
        }

        // BodyXY._get_targvec_img (body_xy.py:3197-3225) incl. the early-out circle
        bool on_disc = false;
        V3 p = mk(nan, nan, nan);
        double lt = 0.0;
        if (mask & kSurfMask) {
            const double dx = x - f.x0, dy = y - f.y0;
            if (!(f.optimize_speed != 0.0 && (dx * dx + dy * dy) > f.r_cut2)) on_disc = sincpt(fs, d, p, lt);
        }

        double v_lon = nan, v_lat = nan, v_clon = nan, v_clat = nan, v_g = nan, v_i = nan, v_e = nan,
               v_az = nan, v_lst = nan, v_dist = nan, v_rv = nan, v_dop = nan;
        if (on_disc) {
            if (mask & kLonLatMask) {  // _get_lonlat_img (body_xy.py:3284-3288)
                double lon, lat, alt;
                recpgr(fs, p, fs.biaxial != 0, lon, lat, alt);
                v_lon = lon * kDpr;
                v_lat = lat * kDpr;
                if (mask & bit(PM_LOCAL_SOLAR_TIME)) v_lst = local_solar_time(f, v_lon);
            }
            if (mask & kCentricMask) {  // _get_lonlat_centric_img (body_xy.py:3349)
                double lon, lat;
                reclat_angles(p, lon, lat);
                v_clon = lon * kDpr;
                v_clat = lat * kDpr;
            }
            if (mask & (kIllumMask | kStateMask | kRingMask)) {
                PointState s;
                if (mask & kIllumMask) {
                    point_state<true, true>(fs, p, lt, s);
                    // _get_illumination_gie_img (body_xy.py:3661-3665)
                    v_g = s.phase * kDpr;
                    v_i = s.incdnc * kDpr;
                    v_e = s.emissn * kDpr;
                    if (mask & bit(PM_AZIMUTH))  // get_azimuth_angle_img (body_xy.py:3744)
                        v_az = azimuth_from_gie(v_g * kRpd, v_i * kRpd, v_e * kRpd) * kDpr;
                } else {
                    point_state<true, false>(fs, p, lt, s);
                }
                v_dist = s.lt * f.clight;  // get_distance_img (body_xy.py:3870-3880)
                v_rv = s.rv;               // get_radial_velocity_img (body_xy.py:3898-3913)
                v_dop = doppler_factor(f, v_rv);
            }
        }
        PM_STORE(PM_LON_GRAPHIC, v_lon);
        PM_STORE(PM_LAT_GRAPHIC, v_lat);
        PM_STORE(PM_LON_CENTRIC, v_clon);
        PM_STORE(PM_LAT_CENTRIC, v_clat);
        PM_STORE(PM_PHASE, v_g);
        PM_STORE(PM_INCIDENCE, v_i);
        PM_STORE(PM_EMISSION, v_e);
        PM_STORE(PM_AZIMUTH, v_az);
        PM_STORE(PM_LOCAL_SOLAR_TIME, v_lst);
        PM_STORE(PM_DISTANCE, v_dist);
        PM_STORE(PM_RADIAL_VELOCITY, v_rv);
        PM_STORE(PM_DOPPLER, v_dop);

        if (mask & kLimbMask) {  // _get_limb_coordinate_imgs (body_xy.py:3967-3975)
            double llon, llat, ldist;
            limb_coordinates(fs, d2, llon, llat, ldist);
            PM_STORE(PM_LIMB_DISTANCE, ldist);
            PM_STORE(PM_LIMB_LON_GRAPHIC, llon);
            PM_STORE(PM_LIMB_LAT_GRAPHIC, llat);
        }
        if (mask & kRingMask) {  // _get_ring_plane_coordinate_imgs (body_xy.py:4061-4085)
            double rad, rl, rd;
            ring_coordinates(fs, d2, rad, rl, rd);
            if (rd > v_dist) rad = rl = rd = nan;  // NaN distance compares false (quirk kept)
            PM_STORE(PM_RING_RADIUS, rad);
            PM_STORE(PM_RING_LON_GRAPHIC, rl);
            PM_STORE(PM_RING_DISTANCE, rd);
        }
    }
}

// ---------------------------------------------------------------------------------
// Map direction: the same quantities on arbitrary planetographic lon/lat cells, plus
// the inverse mapping cell -> image xy with the reference's visibility test.
// Replaces the loops listed at pm_backplanes_map in include/pm_b200.h.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) backplanes_map_kernel(const PMFrame *__restrict__ frame,
                                                                const double *__restrict__ lon_in,
                                                                const double *__restrict__ lat_in,
                                                                int64_t n, uint64_t mask,
                                                                double *__restrict__ out) {
    __shared__ FrameS fs;
    load_frame(fs, frame);
    const PMFrame &f = fs.f;
    const int64_t plane_stride = n;
    const double nan = NAN;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const double lon_deg = __ldg(lon_in + idx), lat_deg = __ldg(lat_in + idx);
        // BodyXY._get_lonlat_map (body_xy.py:3293-3300)
        const double lonm = isfinite(lon_deg) ? pymod_pos(lon_deg, 360.0) : nan;
        const double latm = isfinite(lat_deg) ? lat_deg : nan;
        double v[PM_N_PLANES];
#pragma unroll
        for (int k = 0; k < PM_N_PLANES; k++) v[k] = nan;
        v[PM_LON_GRAPHIC] = lonm;
        v[PM_LAT_GRAPHIC] = latm;
        if (!isnan(lonm)) {
            if (mask & bit(PM_LOCAL_SOLAR_TIME))  // get_local_solar_time_map (body_xy.py:3812)
                v[PM_LOCAL_SOLAR_TIME] = local_solar_time(f, lonm);
            if (isfinite(latm)) {
                V3 tv = pgrrec0(fs, lonm * kRpd, latm * kRpd);  // _get_targvec_map (:3230)
                if (mask & kCentricMask) {                        // _get_lonlat_centric_map (:3357)
                    double lo, la;
                    reclat_angles(tv, lo, la);
                    v[PM_LON_CENTRIC] = lo * kDpr;
                    v[PM_LAT_CENTRIC] = la * kDpr;
                }
                PointState s;
                point_state<true, true>(fs, tv, f.lt0, s);  // _get_illumf_map (:3671), _get_state_maps (:3851)
                const double g = s.phase * kDpr, i = s.incdnc * kDpr, e = s.emissn * kDpr;
                v[PM_PHASE] = g;
                v[PM_INCIDENCE] = i;
                v[PM_EMISSION] = e;
                if (mask & bit(PM_AZIMUTH)) v[PM_AZIMUTH] = azimuth_from_gie(g * kRpd, i * kRpd, e * kRpd) * kDpr;
                const double dist = s.lt * f.clight;
                v[PM_DISTANCE] = dist;
                v[PM_RADIAL_VELOCITY] = s.rv;
                v[PM_DOPPLER] = doppler_factor(f, s.rv);
                const bool visibl = s.emissn < kHalfPi, lit = s.incdnc < kHalfPi;
                const uint64_t vis_mask = bit(PM_RA) | bit(PM_DEC) | bit(PM_PIXEL_X) | bit(PM_PIXEL_Y) | kKmMask;
                V3 ov = mk(nan, nan, nan);
                if ((visibl && (mask & vis_mask)) || (lit && (mask & (kLimbMask | kRingMask))))
                    ov = targvec2obsvec(fs, tv);  // _get_obsvec_map (:3275)
                if (visibl && (mask & vis_mask)) {
                    double ra, dec;  // _get_radec_map (:3423-3432)
                    recrad_angles(ov, ra, dec);
                    const double ra_deg = ra * kDpr, dec_deg = dec * kDpr;
                    v[PM_RA] = ra_deg;
                    v[PM_DEC] = dec_deg;
                    V3 d2 = radrec1(ra_deg * kRpd, dec_deg * kRpd);
                    double ax, ay;
                    obsvec2angular(f, d2, ax, ay);
                    // _get_xy_map (:3482-3491) with _xy_in_image_frame (:1868)
                    const double x = f.Ainv[0] * ax + f.Ainv[1] * ay + f.Ainv[2];
                    const double y = f.Ainv[3] * ax + f.Ainv[4] * ay + f.Ainv[5];
                    if ((-0.5 < x && x < f.nx - 0.5) && (-0.5 < y && y < f.ny - 0.5)) {
                        v[PM_PIXEL_X] = x;
                        v[PM_PIXEL_Y] = y;
                    }
                    const double kx = f.ang2km[0] * ax + f.ang2km[1] * ay;  // _get_km_xy_map (:3557)
                    const double ky = f.ang2km[2] * ax + f.ang2km[3] * ay;
                    v[PM_KM_X] = kx;
                    v[PM_KM_Y] = ky;
                    v[PM_ANGULAR_X] = kx / f.km_per_arcsec;
                    v[PM_ANGULAR_Y] = ky / f.km_per_arcsec;
                }
                // the reference tests `lit` (illumf[4]) here, not `visibl`
                // (body_xy.py:3981, :4097); reproduced as is
                if (lit && (mask & kLimbMask))
                    limb_coordinates(fs, ov, v[PM_LIMB_LON_GRAPHIC], v[PM_LIMB_LAT_GRAPHIC], v[PM_LIMB_DISTANCE]);
                if (lit && (mask & kRingMask)) {
                    double rad, rl, rd;
                    ring_coordinates(fs, ov, rad, rl, rd);
                    if (rd > dist) rad = rl = rd = nan;
                    v[PM_RING_RADIUS] = rad;
                    v[PM_RING_LON_GRAPHIC] = rl;
                    v[PM_RING_DISTANCE] = rd;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < PM_N_PLANES; k++) PM_STORE(k, v[k]);
    }
}

// BodyXY._xy2lonlat (body_xy.py:482-496) -> Body._obsvec_norm2lonlat (body.py:1058-1081)
__global__ void __launch_bounds__(kBlock) xy2lonlat_kernel(const PMFrame *__restrict__ frame,
                                                           const double *__restrict__ xs,
                                                           const double *__restrict__ ys, int64_t n,
                                                           double *__restrict__ lon_out,
                                                           double *__restrict__ lat_out,
                                                           unsigned long long *__restrict__ n_missed) {
    __shared__ FrameS fs;
    load_frame(fs, frame);
    unsigned long long missed = 0;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const double x = xs[idx], y = ys[idx];
        double lon = NAN, lat = NAN;
        if (isfinite(x) && isfinite(y)) {
            V3 d = xy2obsvec_norm(fs.f, x, y);
            V3 p;
            double lt;
            if (sincpt(fs, d, p, lt)) {
                double lo, la, al;
                recpgr(fs, p, fs.biaxial != 0, lo, la, al);
                lon = lo * kDpr;
                lat = la * kDpr;
            } else {
                missed++;
            }
        }
        lon_out[idx] = lon;
        lat_out[idx] = lat;
    }
    if (n_missed) {
        for (int o = 16; o > 0; o >>= 1) missed += __shfl_down_sync(0xffffffffu, missed, o);
        if ((threadIdx.x & 31) == 0 && missed) atomicAdd(n_missed, missed);
    }
}

// BodyXY._lonlat2xy (body_xy.py:544-560) -> Body._lonlat2obsvec (body.py:1039-1056),
// alt == 0 visibility via illumf.visibl (body.py:2124-2130)
__global__ void __launch_bounds__(kBlock) lonlat2xy_kernel(const PMFrame *__restrict__ frame,
                                                           const double *__restrict__ lons,
                                                           const double *__restrict__ lats, int64_t n,
                                                           uint32_t flags, double *__restrict__ x_out,
                                                           double *__restrict__ y_out) {
    __shared__ FrameS fs;
    load_frame(fs, frame);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const double lon = lons[idx], lat = lats[idx];
        double x = NAN, y = NAN;
        if (isfinite(lon) && isfinite(lat)) {
            V3 tv = pgrrec0(fs, lon * kRpd, lat * kRpd);
            bool keep = true;
            if (flags & PM_FLAG_NOT_VISIBLE_NAN) {
                PointState s;
                point_state<false, true>(fs, tv, fs.f.lt0, s);
                keep = s.emissn < kHalfPi;
            }
            if (keep) obsvec2xy(fs.f, targvec2obsvec(fs, tv), x, y);
        }
        x_out[idx] = x;
        y_out[idx] = y;
    }
}

// FP64 FMA throughput probe (roofline denominator for the compute-bound kernels)
__global__ void __launch_bounds__(256) fp64_probe_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0;
    double a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 0.999999999, c = 1e-12;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) out[0] = s;
}

// ---------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------
static int grid_for(int64_t n, int sm_count) {
    int64_t blocks = (n + kBlock - 1) / kBlock;
    int64_t cap = (int64_t)sm_count * 32;  // a few waves of resident CTAs; kernels grid-stride
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

cudaError_t launch_backplanes_img(const PMFrame *frames, int n_frames, int nx, int ny, uint64_t mask,
                                  double *out, int sm_count, cudaStream_t st) {
    dim3 grid(grid_for((int64_t)nx * ny, sm_count), n_frames);
    backplanes_img_kernel<<<grid, kBlock, 0, st>>>(frames, nx, ny, mask, out);
    count_launches(1);
    return cudaGetLastError();
}
cudaError_t launch_backplanes_map(const PMFrame *frame, const double *lon, const double *lat, int64_t n,
                                  uint64_t mask, double *out, int sm_count, cudaStream_t st) {
    backplanes_map_kernel<<<grid_for(n, sm_count), kBlock, 0, st>>>(frame, lon, lat, n, mask, out);
    count_launches(1);
    return cudaGetLastError();
}
cudaError_t launch_xy2lonlat(const PMFrame *frame, const double *x, const double *y, int64_t n, double *lon,
                             double *lat, unsigned long long *n_missed, int sm_count, cudaStream_t st) {
    xy2lonlat_kernel<<<grid_for(n, sm_count), kBlock, 0, st>>>(frame, x, y, n, lon, lat, n_missed);
    count_launches(1);
    return cudaGetLastError();
}
cudaError_t launch_lonlat2xy(const PMFrame *frame, const double *lon, const double *lat, int64_t n,
                             uint32_t flags, double *x, double *y, int sm_count, cudaStream_t st) {
    lonlat2xy_kernel<<<grid_for(n, sm_count), kBlock, 0, st>>>(frame, lon, lat, n, flags, x, y);
    count_launches(1);
    return cudaGetLastError();
}
cudaError_t launch_fp64_probe(double *scratch, int iters, int sm_count, cudaStream_t st) {
    fp64_probe_kernel<<<sm_count * 8, 256, 0, st>>>(scratch, iters);
    count_launches(1);
    return cudaGetLastError();
}

}  // namespace pm
