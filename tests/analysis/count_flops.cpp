// count_flops.cpp - derives the ALGORITHMIC flop count per pixel of the image
// backplane path by running the CPU oracle (oracle/pm_oracle.c) with `double`
// replaced by a counting scalar type (SURVEY.md section 8(d) asks for exactly this).
//
//   g++ -O1 -o /tmp/count_flops tools/count_flops.cpp && /tmp/count_flops frame.bin nx ny mask
//
// Weights (approximate FP64 instruction costs): add/sub/mul = 1, div = 18,
// sqrt = 15, sin or cos = 45 each (sincos pair = 90), atan2 = 80, asin/acos = 70,
// fmod = 20, floor / nearbyint / fmax / fmin = 1; comparisons, fabs, negation = 0.
// Output: JSON with the mean weighted flops of the three pixel classes.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static unsigned long long g_w = 0;  // weighted flops
static unsigned long long g_n[8] = {0};  // raw: 0 add/sub/mul, 1 div, 2 sqrt, 3 sin/cos, 4 atan2, 5 asin/acos, 6 other

struct C {
    double v;
    C() : v(0) {}
    C(double x) : v(x) {}
    C(int x) : v(x) {}
    C(long x) : v((double)x) {}
    explicit operator double() const { return v; }
    explicit operator long() const { return (long)v; }
    explicit operator int() const { return (int)v; }
};
#define BINOP(op, w, slot)                                                        \
    static inline C operator op(C a, C b) { g_w += w; g_n[slot]++; return C(a.v op b.v); }         \
    static inline C operator op(C a, double b) { g_w += w; g_n[slot]++; return C(a.v op b); }      \
    static inline C operator op(double a, C b) { g_w += w; g_n[slot]++; return C(a op b.v); }      \
    static inline C operator op(C a, int b) { g_w += w; g_n[slot]++; return C(a.v op b); }         \
    static inline C operator op(int a, C b) { g_w += w; g_n[slot]++; return C(a op b.v); }
BINOP(+, 1, 0)
BINOP(-, 1, 0)
BINOP(*, 1, 0)
BINOP(/, 18, 1)
static inline C operator-(C a) { return C(-a.v); }
static inline C &operator+=(C &a, C b) { g_w += 1; g_n[0]++; a.v += b.v; return a; }
static inline C &operator-=(C &a, C b) { g_w += 1; g_n[0]++; a.v -= b.v; return a; }
static inline C &operator*=(C &a, C b) { g_w += 1; g_n[0]++; a.v *= b.v; return a; }
static inline C &operator/=(C &a, C b) { g_w += 18; g_n[1]++; a.v /= b.v; return a; }
static inline C &operator+=(C &a, double b) { g_w += 1; g_n[0]++; a.v += b; return a; }
static inline C &operator-=(C &a, double b) { g_w += 1; g_n[0]++; a.v -= b; return a; }
#define CMP(op)                                                      \
    static inline bool operator op(C a, C b) { return a.v op b.v; }  \
    static inline bool operator op(C a, double b) { return a.v op b; } \
    static inline bool operator op(double a, C b) { return a op b.v; } \
    static inline bool operator op(C a, int b) { return a.v op b; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
#define FN1(name, w, slot) static inline C name(C a) { g_w += w; g_n[slot]++; return C(std::name(a.v)); }
FN1(sqrt, 15, 2) FN1(sin, 45, 3) FN1(cos, 45, 3) FN1(asin, 70, 5) FN1(acos, 70, 5) FN1(floor, 1, 6)
FN1(nearbyint, 1, 6)
static inline C fabs(C a) { return C(std::fabs(a.v)); }
static inline C atan2(C a, C b) { g_w += 80; g_n[4]++; return C(std::atan2(a.v, b.v)); }
static inline C fmod(C a, C b) { g_w += 20; g_n[6]++; return C(std::fmod(a.v, b.v)); }
static inline C fmod(C a, double b) { g_w += 20; g_n[6]++; return C(std::fmod(a.v, b)); }
static inline C hypot(C a, C b) { g_w += 18; g_n[2]++; return C(std::hypot(a.v, b.v)); }
static inline C fmax(C a, C b) { g_w += 1; g_n[6]++; return C(std::fmax(a.v, b.v)); }
static inline C fmax(double a, C b) { g_w += 1; g_n[6]++; return C(std::fmax(a, b.v)); }
static inline C fmax(C a, double b) { g_w += 1; g_n[6]++; return C(std::fmax(a.v, b)); }
static inline C fmin(C a, C b) { g_w += 1; g_n[6]++; return C(std::fmin(a.v, b.v)); }
static inline C fmin(double a, C b) { g_w += 1; g_n[6]++; return C(std::fmin(a, b.v)); }
static inline C fmin(C a, double b) { g_w += 1; g_n[6]++; return C(std::fmin(a.v, b)); }
static inline bool isfinite(C a) { return std::isfinite(a.v); }
static inline bool isnan(C a) { return std::isnan(a.v); }

#define double C
#include "../../oracle/pm_oracle.c"
#undef double

int main(int argc, char **argv) {
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s frame.bin nx ny mask\n", argv[0]);
        return 2;
    }
    PMFrame f;
    FILE *fp = std::fopen(argv[1], "rb");
    if (!fp || std::fread(&f, sizeof(PMFrame), 1, fp) != 1) return 3;
    std::fclose(fp);
    int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]);
    uint64_t mask = std::strtoull(argv[4], nullptr, 0);
    unsigned long long w[3] = {0, 0, 0}, n[3] = {0, 0, 0}, raw[3][8] = {{0}};
    for (int y = 0; y < ny; y++)
        for (int x = 0; x < nx; x++) {
            unsigned long long w0 = g_w, r0[8];
            std::memcpy(r0, g_n, sizeof(r0));
            PixelOut o;
            pixel_backplanes(&f, C((double)x), C((double)y), mask, &o);
            ::C dx = (double)x - (double)f.x0, dy = (double)y - (double)f.y0;
            bool outside = (double)(dx.v * dx.v + dy.v * dy.v) > (double)f.r_cut2;
            bool on = !std::isnan((double)o.margin) && std::isfinite((double)o.v[PM_DISTANCE]);
            int cls = outside ? 2 : (on ? 0 : 1);
            if ((mask & (1ull << PM_DISTANCE)) == 0) cls = outside ? 2 : (std::isfinite((double)o.v[PM_LON_GRAPHIC]) ? 0 : 1);
            w[cls] += g_w - w0;
            n[cls]++;
            for (int k = 0; k < 8; k++) raw[cls][k] += g_n[k] - r0[k];
        }
    const char *names[3] = {"on_disc", "in_circle_miss", "outside_circle"};
    std::printf("{");
    for (int c = 0; c < 3; c++) {
        std::printf("\"%s\": %.2f, \"%s_px\": %llu, \"%s_raw\": [", names[c], n[c] ? (double)w[c] / n[c] : 0.0,
                    names[c], n[c], names[c]);
        for (int k = 0; k < 7; k++) std::printf("%.2f%s", n[c] ? (double)raw[c][k] / n[c] : 0.0, k < 6 ? ", " : "");
        std::printf("]%s", c < 2 ? ", " : "");
    }
    std::printf("}\n");
    return 0;
}
