// Separable RoIAlign axis arithmetic, shared by the CUDA kernels (device) and the host
// emulation used by the CPU tests (tests/emu).  No CUDA headers needed on the host side.
//
// Reference semantics (mmcv.ops.RoIAlign, aligned=True, avg pool; the reference constructs it
// at mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55 and calls it at
// single_level_roi_extractor.py:93 and adaptative_roi_extractor.py:72,87):
//   sample coordinate  t(p,i) = start + p*bin + (i+0.5)*bin/g ,  g = ceil(roi_len/P)
//   a sample is dropped when t < -1 or t > L; otherwise t is clamped to >= 0,
//   lo = (int)t, and lo >= L-1 collapses both taps onto L-1.
// Because t_y depends only on (ph,iy) and t_x only on (pw,ix), and the validity test and the
// bilinear weights factorise, the averaged output of one channel is
//   Y[ph][pw] = sum_r sum_c  Wy[ph][r] * Wx[pw][c] * F[r][c]
// with per-axis weights  W[p][j] = (1/g) * sum_i tapweight(t(p,i) -> pixel j).
// All coordinate arithmetic is fp64 with explicit round-to-nearest ops (no fma contraction),
// evaluated in the same order as the fp64 run of the reference, so weights agree with the fp64
// oracle to the last bit before the final rounding to fp32.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define HTD_HD __host__ __device__ __forceinline__
#else
#define HTD_HD inline
#endif

namespace htd {

HTD_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
HTD_HD double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
HTD_HD double ddiv(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}

struct Axis {
    double start;  // roi start in feature pixels: lo*scale - 0.5
    double bin;    // roi_len / P
    double dt;     // bin / g  (spacing of consecutive samples), 0 when all samples coincide
    int g;         // samples per bin; <= 0 -> the bin has no samples (output 0)
    int L;         // feature-map extent along this axis
};

// lo/hi are the fp32 roi corners (image coordinates), scale = 1/stride.
HTD_HD Axis make_axis(float lo, float hi, double scale, int P, int L, int sampling_ratio,
                      int aligned) {
    Axis a;
    double off = aligned ? 0.5 : 0.0;
    double s = dadd(dmul((double)lo, scale), -off);
    double e = dadd(dmul((double)hi, scale), -off);
    double len = dadd(e, -s);
    if (!aligned && len < 1.0) len = 1.0;
    a.start = s;
    a.bin = ddiv(len, (double)P);
    a.L = L;
    a.g = 0;
    a.dt = 0.0;
    // NaN / negative / absurd extents produce no samples (for len <= 0 and sampling_ratio == 0
    // this IS the reference result: ceil(len/P) <= 0; for sampling_ratio > 0 with len < 0 the
    // reference samples a reversed grid - documented deviation, never produced by the head).
    if (!(len >= 0.0) || !(len < 1e9)) return a;
    if (sampling_ratio > 0) {
        if (len == 0.0) { a.g = 1; a.dt = 0.0; }          // all samples coincide: one tap set
        else { a.g = sampling_ratio; a.dt = ddiv(a.bin, (double)a.g); }
    } else {
        a.g = (int)ceil(a.bin);
        if (a.g > 0) a.dt = ddiv(a.bin, (double)a.g);
    }
    return a;
}

// same expression order as the oracle: (start + p*bin) + ((i + 0.5)*bin)/g
HTD_HD double sample_t(const Axis& a, int p, int i) {
    return dadd(dadd(a.start, dmul((double)p, a.bin)),
                ddiv(dmul(dadd((double)i, 0.5), a.bin), (double)a.g));
}

HTD_HD bool tap(const Axis& a, double t, int& lo, int& hi, double& frac) {
    if (!(t >= -1.0 && t <= (double)a.L)) return false;
    if (t <= 0.0) t = 0.0;
    lo = (int)t;
    if (lo >= a.L - 1) { lo = hi = a.L - 1; frac = 0.0; }
    else { hi = lo + 1; frac = t - (double)lo; }
    return true;
}

HTD_HD int clampi_from_double(double x, int lo, int hi) {
    if (!(x > (double)lo)) return lo;
    if (!(x < (double)hi)) return hi;
    return (int)x;
}

// Inclusive pixel range [jlo, jhi] touched by bin p; empty when jhi < jlo.
HTD_HD void bin_range(const Axis& a, int p, int& jlo, int& jhi) {
    jlo = 0; jhi = -1;
    if (a.g <= 0) return;
    int ifirst = 0, ilast = a.g - 1;
    if (a.dt > 0.0) {
        double t0 = sample_t(a, p, 0);
        // smallest i with t_i >= -1, largest i with t_i <= L (t increases with i)
        ifirst = clampi_from_double(ceil((-1.0 - t0) / a.dt), 0, a.g);
        while (ifirst > 0 && sample_t(a, p, ifirst - 1) >= -1.0) --ifirst;
        while (ifirst < a.g && !(sample_t(a, p, ifirst) >= -1.0)) ++ifirst;
        ilast = clampi_from_double(floor(((double)a.L - t0) / a.dt), -1, a.g - 1);
        while (ilast < a.g - 1 && sample_t(a, p, ilast + 1) <= (double)a.L) ++ilast;
        while (ilast >= 0 && !(sample_t(a, p, ilast) <= (double)a.L)) --ilast;
    }
    if (ifirst > ilast) return;
    int lo, hi; double fr;
    if (!tap(a, sample_t(a, p, ifirst), lo, hi, fr)) return;
    int first_lo = lo;
    if (!tap(a, sample_t(a, p, ilast), lo, hi, fr)) return;
    jlo = first_lo; jhi = hi;
}

// Weight of feature pixel j for output bin p (already divided by g).
HTD_HD float axis_weight(const Axis& a, int p, int j) {
    if (a.g <= 0 || j < 0 || j >= a.L) return 0.f;
    int ilo = 0, ihi = a.g - 1;
    if (a.dt > 0.0) {
        // samples that can touch pixel j lie in (j-1, j+1); the clamping at the borders widens
        // the window to [-1, 1) for j == 0 and (L-2, L] for j == L-1
        double tlo = (j == 0) ? -1.0 : (double)(j - 1);
        double thi = (j == a.L - 1) ? (double)a.L : (double)(j + 1);
        double t0 = sample_t(a, p, 0);
        ilo = clampi_from_double(floor((tlo - t0) / a.dt) - 1.0, 0, a.g);
        ihi = clampi_from_double(ceil((thi - t0) / a.dt) + 1.0, -1, a.g - 1);
    }
    double acc = 0.0;
    for (int i = ilo; i <= ihi; ++i) {
        int lo, hi; double fr;
        if (!tap(a, sample_t(a, p, i), lo, hi, fr)) continue;
        if (lo == j) acc += 1.0 - fr;
        if (hi == j && hi != lo) acc += fr;
    }
    return (float)(acc / (double)a.g);
}

// Union of the pixel ranges of all P bins (the RoI footprint along this axis).
HTD_HD void roi_range(const Axis& a, int P, int& jlo, int& jhi) {
    jlo = 0; jhi = -1;
    bool any = false;
    for (int p = 0; p < P; ++p) {
        int lo, hi;
        bin_range(a, p, lo, hi);
        if (hi < lo) continue;
        if (!any) { jlo = lo; jhi = hi; any = true; }
        else { if (lo < jlo) jlo = lo; if (hi > jhi) jhi = hi; }
    }
}

// FPN level assignment: clamp(floor(log2(sqrt(w*h)/finest + 1e-6)), 0, L-1), fp32 semantics of
// SingleRoIExtractor.map_roi_levels (single_level_roi_extractor.py:47-50).  log2 is evaluated in
// fp64 and rounded once to fp32, i.e. a correctly rounded log2f, which is what glibc / torch-CPU
// produce on the golden boundary RoIs (tests/golden/levels.npz).  NaN scale -> -1 ("no level";
// the reference's .long() of NaN matches no level either).
HTD_HD int roi_level(float x1, float y1, float x2, float y2, float finest_scale, int num_levels) {
#if defined(__CUDA_ARCH__)
    float area = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
    float s = __fsqrt_rn(area);
    float v = __fadd_rn(__fdiv_rn(s, finest_scale), 1e-6f);
#else
    float area = (x2 - x1) * (y2 - y1);
    float s = sqrtf(area);
    float v = s / finest_scale + 1e-6f;
#endif
    if (!(v == v)) return -1;
    float t = floorf((float)log2((double)v));
    if (t < 0.f) t = 0.f;
    if (t > (float)(num_levels - 1)) t = (float)(num_levels - 1);
    return (int)t;
}

}  // namespace htd
