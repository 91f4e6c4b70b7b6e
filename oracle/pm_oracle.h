/*
 * pm_oracle.h - entry points of the CPU oracle (test infrastructure only; see the
 * header of pm_oracle.c).  Same PMFrame and plane ids as include/pm_b200.h; all
 * pointers are HOST pointers.  `margin` (may be NULL) receives the tangency margin
 * used by the parity tests to exclude and count grazing pixels: image direction =
 * |perpendicular ray offset| - 1 in the unit-sphere-scaled frame (0 at the limb),
 * map direction = emission - pi/2.
 */
#ifndef PM_ORACLE_H
#define PM_ORACLE_H
#include "../include/pm_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
int pmo_backplanes_img(const PMFrame *f, int nx, int ny, uint64_t mask, double *out,
                       double *margin);
int pmo_backplanes_map(const PMFrame *f, const double *lon, const double *lat, int64_t n,
                       uint64_t mask, double *out, double *margin);
int pmo_xy2lonlat(const PMFrame *f, const double *x, const double *y, int64_t n,
                  double *lon, double *lat, int64_t *n_missed);
int pmo_lonlat2xy(const PMFrame *f, const double *lon, const double *lat, int64_t n,
                  uint32_t flags, double *x, double *y);
int pmo_lonlat2xy_alt(const PMFrame *f, const double *lon, const double *lat, int64_t n, double alt,
                      uint32_t flags, double *x, double *y);
int pmo_transform(const PMFrame *f, int src, int dst, const double *a, const double *b, int64_t n, double alt,
                  uint32_t flags, const double *aux13, double *out_a, double *out_b, int64_t *n_missed);
int pmo_proj_inverse(int kind, const double *params5, const double *xx, const double *yy,
                     int64_t n, double *lon, double *lat);
int pmo_proj_forward(int kind, const double *params5, const double *lon, const double *lat, int64_t n,
                     double *xx, double *yy);
int pmo_gather_nearest(const double *cube, int n_planes, int ny, int nx,
                       const double *xmap, const double *ymap, int64_t n_cells,
                       double *out);
#ifdef __cplusplus
}
#endif
#endif
