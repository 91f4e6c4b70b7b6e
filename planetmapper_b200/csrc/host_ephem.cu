// host_ephem.cu - HOST code (no device functions): the two ephemeris primitives the once-per-frame
// constants are built from, for providers without spiceypy.
//
// In the reference these are CSPICE calls made through spiceypy (spkssb / spkezr at
// planetmapper/base.py:828, pxform / sxform behind body.py:935-945); the Python MiniSpice reader
// (planetmapper_b200/minispice) restates them from the public NAIF "SPK Required Reading" (types 2
// and 3: Chebyshev records MID, RADIUS, coefficients) and the IAU orientation model of the text PCK
// (pole RA / Dec and prime meridian polynomials plus nutation-precession series).  A frame needs
// about thirty such evaluations; in Python they are ~80 % of its ~1 ms cost, which is what bounds
// a time series once the kernels take 40 microseconds per frame.  These functions evaluate the
// same formulas in the same order on the same tables (planetmapper_b200/minispice/native.py hands
// them over), so results agree with the Python reader to the last bit or two.
#include <math.h>
#include <stdint.h>

#include "../../include/pm_b200.h"

namespace {

constexpr double kSpd = 86400.0;
constexpr double kPiH = 3.14159265358979323846264338327950288;

struct Mat3 {
    double m[3][3];
};
Mat3 mul(const Mat3 &a, const Mat3 &b) {
    Mat3 c;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
            c.m[i][j] = s;
        }
    return c;
}
Mat3 scale(const Mat3 &a, double f) {
    Mat3 c;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) c.m[i][j] = a.m[i][j] * f;
    return c;
}
Mat3 add3(const Mat3 &a, const Mat3 &b, const Mat3 &c) {
    Mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = (a.m[i][j] + b.m[i][j]) + c.m[i][j];
    return r;
}
// spice.rotate: coordinate-system rotation by theta about axis 1 / 2 / 3, and its derivative
Mat3 rot_axis(double theta, int axis, bool derivative) {
    const double c = cos(theta), s = sin(theta);
    Mat3 r = {{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}};
    const double one = derivative ? 0.0 : 1.0;
    const double cc = derivative ? -s : c, ss = derivative ? c : s;  // d/dtheta (c, s) = (-s, c)
    if (axis == 1) {
        r.m[0][0] = one;
        r.m[1][1] = cc;
        r.m[1][2] = ss;
        r.m[2][1] = -ss;
        r.m[2][2] = cc;
    } else if (axis == 2) {
        r.m[1][1] = one;
        r.m[0][0] = cc;
        r.m[0][2] = -ss;
        r.m[2][0] = ss;
        r.m[2][2] = cc;
    } else {
        r.m[2][2] = one;
        r.m[0][0] = cc;
        r.m[0][1] = ss;
        r.m[1][0] = -ss;
        r.m[1][1] = cc;
    }
    return r;
}

double poly(const double *c, int n, double t) {  // c0 + c1 t + c2 t^2, as the Python reader writes it
    double v = n > 0 ? c[0] : 0.0;
    if (n > 1) v += c[1] * t;
    if (n > 2) v += c[2] * t * t;
    return v;
}
double dpoly(const double *c, int n, double t) {
    double v = n > 1 ? c[1] : 0.0;
    if (n > 2) v += 2 * c[2] * t;
    return v;
}

// record index of a segment at et, or -1 if the segment (or the kept part of it) does not cover et
int64_t record_of(const PMEphemSegment &s, double et) {
    if (!(s.et_start <= et && et <= s.et_end)) return -1;
    int64_t idx = (int64_t)floor((et - s.init) / s.intlen);
    if (idx < 0) idx = 0;
    if (idx > s.n - 1) idx = s.n - 1;
    if (idx < s.first_record || idx >= (int64_t)s.first_record + s.n_kept) return -1;
    return idx - s.first_record;
}

}  // namespace

extern "C" {

int pm_host_ssb_state(const PMEphemSegment *segs, int n_segs, const double *records, int body, double et,
                      double *state6) {
    if (!segs || !records || !state6 || n_segs < 0) return PM_ERR_BAD_ARG;
    double acc[6] = {0, 0, 0, 0, 0, 0};
    int b = body;
    for (int guard = 0; b != 0; guard++) {
        if (guard > 8) return PM_ERR_BAD_ARG;  // centre chain does not reach the barycentre
        const PMEphemSegment *seg = nullptr;
        int64_t ridx = -1;
        for (int i = 0; i < n_segs; i++) {  // highest precedence first
            if (segs[i].target != b) continue;
            ridx = record_of(segs[i], et);
            if (ridx >= 0) {
                seg = &segs[i];
                break;
            }
        }
        if (!seg) return PM_ERR_UNSUPPORTED;  // no SPK data for this body at et
        const double *rec = records + seg->rec_offset + ridx * seg->rsize;
        const double mid = rec[0], radius = rec[1];
        const double s = (et - mid) / radius, two_s = 2.0 * s;
        const int rows = seg->spk_type == 2 ? 3 : 6;
        const int ncoef = (seg->rsize - 2) / rows;
        if (ncoef > 64 || (seg->spk_type != 2 && seg->spk_type != 3)) return PM_ERR_UNSUPPORTED;
        double t[64], dt[64];
        t[0] = 1.0;
        dt[0] = 0.0;
        if (ncoef > 1) {
            t[1] = s;
            dt[1] = 1.0;
        }
        for (int k = 2; k < ncoef; k++) {
            t[k] = two_s * t[k - 1] - t[k - 2];
            dt[k] = 2.0 * t[k - 1] + two_s * dt[k - 1] - dt[k - 2];
        }
        const double *coefs = rec + 2;
        double val[6] = {0, 0, 0, 0, 0, 0};
        for (int r = 0; r < rows; r++) {
            double v = 0.0;
            for (int k = 0; k < ncoef; k++) v += coefs[r * ncoef + k] * t[k];
            val[r] = v;
        }
        if (seg->spk_type == 2) {  // velocity by differentiating the position polynomial
            for (int r = 0; r < 3; r++) {
                double v = 0.0;
                for (int k = 0; k < ncoef; k++) v += coefs[r * ncoef + k] * dt[k];
                val[3 + r] = v / radius;
            }
        }
        for (int i = 0; i < 6; i++) acc[i] += val[i];
        b = seg->center;
    }
    for (int i = 0; i < 6; i++) state6[i] = acc[i];
    return PM_OK;
}

int pm_host_orientation(const PMOrientationModel *m, double et, double *rmat9, double *omega3) {
    if (!m || !rmat9 || !omega3 || m->n_nut < 0 || m->n_nut > PM_MAX_NUT_TERMS) return PM_ERR_BAD_ARG;
    const double century = kSpd * 36525.0;
    const double t_cy = et / century, d = et / kSpd;
    double ra = poly(m->pole_ra, m->n_ra, t_cy), dra = dpoly(m->pole_ra, m->n_ra, t_cy) / century;
    double dec = poly(m->pole_dec, m->n_dec, t_cy), ddec = dpoly(m->pole_dec, m->n_dec, t_cy) / century;
    double w = poly(m->pm, m->n_pm, d), dw = dpoly(m->pm, m->n_pm, d) / kSpd;
    const double rad = kPiH / 180.0;
    if (m->n_nut_ra > 0 || m->n_nut_dec > 0 || m->n_nut_pm > 0) {
        double sn[PM_MAX_NUT_TERMS], cs[PM_MAX_NUT_TERMS], dth[PM_MAX_NUT_TERMS];
        for (int k = 0; k < m->n_nut; k++) {
            const double theta = (m->nut_angles[2 * k] + m->nut_angles[2 * k + 1] * t_cy) * rad;
            sn[k] = sin(theta);
            cs[k] = cos(theta);
            dth[k] = (m->nut_angles[2 * k + 1] * rad) / century;
        }
        double a = 0.0, b = 0.0;
        for (int k = 0; k < m->n_nut_ra; k++) {
            a += m->nut_ra[k] * sn[k];
            b += m->nut_ra[k] * (cs[k] * dth[k]);
        }
        ra += a;
        dra += b;
        a = b = 0.0;
        for (int k = 0; k < m->n_nut_dec; k++) {
            a += m->nut_dec[k] * cs[k];
            b += -m->nut_dec[k] * (sn[k] * dth[k]);
        }
        dec += a;
        ddec += b;
        a = b = 0.0;
        for (int k = 0; k < m->n_nut_pm; k++) {
            a += m->nut_pm[k] * sn[k];
            b += m->nut_pm[k] * (cs[k] * dth[k]);
        }
        w += a;
        dw += b;
    }
    w = fmod(w, 360.0);
    ra *= rad;
    dec *= rad;
    w *= rad;
    dra *= rad;
    ddec *= rad;
    dw *= rad;
    const double a3 = w, a1 = kPiH / 2 - dec, b3 = kPiH / 2 + ra;
    const Mat3 r3 = rot_axis(a3, 3, false), r1 = rot_axis(a1, 1, false), q3 = rot_axis(b3, 3, false);
    const Mat3 rmat = mul(mul(r3, r1), q3);
    const Mat3 drmat = add3(mul(mul(scale(rot_axis(a3, 3, true), dw), r1), q3),
                            mul(mul(r3, scale(rot_axis(a1, 1, true), -ddec)), q3),
                            mul(mul(r3, r1), scale(rot_axis(b3, 3, true), dra)));
    // om = -drmat rmat^T; omega = (om[2][1], om[0][2], om[1][0])
    auto om = [&](int i, int j) {
        double s = 0.0;
        for (int k = 0; k < 3; k++) s += drmat.m[i][k] * rmat.m[j][k];
        return -s;
    };
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) rmat9[3 * i + j] = rmat.m[i][j];
    omega3[0] = om(2, 1);
    omega3[1] = om(0, 2);
    omega3[2] = om(1, 0);
    return PM_OK;
}

}  // extern "C"
