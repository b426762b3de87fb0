// pm_device.cuh -- device-side building blocks shared by the sm_100a kernels.
//
// Everything here is FP64: the reference evaluates on OpenMM's double-precision Reference platform
// (ParaMol/Objective_function/objective_function.py:81 forces OPENMM_DEFAULT_PLATFORM=Reference) and the
// parity bar (objective 1e-8 relative on as few as 10 conformations) cannot be met with FP32 terms.
// B200 issues DFMA at half the FFMA rate, so the design goal is to keep the FP64 pipe saturated and to
// spend as few FP64 instructions per interaction as possible (no libm sqrt / div / acos / sincos in the hot loops).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PM_IT 4          // atoms of an i-tile kept in registers in the pair loop
#ifndef PM_MAX_WARPS
#define PM_MAX_WARPS 16  // warps per CTA (1 persistent CTA per SM)
#endif
#define PM_PI 3.14159265358979323846
// The table behind pm_acos_cs: tabulated angles theta_e, e = 0 .. PM_ACOS_ENTRIES-1, two double2 per entry: (cos, sin)(theta_e) and
// (theta_e, 0).  Entries are indexed by round(K cos theta) where |cos theta| <= 1/sqrt(2) (2 NA + 1 of them) and by round(K sin theta)
// elsewhere (NA + 1 for cos > 0, NA + 1 for cos < 0), so that the nearest entry is found without an inverse trigonometric function
// and is never further than 1.42 / (2 K) = 0.011 rad away.  PM_ACOS_N + 1 = number of double2 the kernels stage into shared memory.
#define PM_ACOS_K 64
#define PM_ACOS_NA 46
#define PM_ACOS_ENTRIES (4 * PM_ACOS_NA + 3)
#define PM_ACOS_N (2 * PM_ACOS_ENTRIES - 1)

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA 1-D bulk copy (SASS: UBLKCP / SYNCS)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pm_mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pm_fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void pm_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void pm_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(pm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     pm_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(pm_smem_u32(bar))
                 : "memory");
}
// L2 prefetch of a contiguous block (one instruction): next tile's coordinates / this tile's reference forces
__device__ __forceinline__ void pm_bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pm_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool pm_mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(pm_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void pm_mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!pm_mbar_try_wait(bar, parity)) {
    }
}

// ---------------------------------------------------------------------------------------------
// FP64 reciprocal square root: MUFU.RSQ64H seed (upper 32 bits, ~2^-20) + ONE cubically convergent
// correction  y (1 + e/2 + 3 e^2 / 8),  e = 1 - x y^2   ->  relative error ~ (5/16) e^3 < 2^-56.
// 1 MUFU + 5 FP64 ops instead of the ~25-instruction libm sqrt + division.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pm_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = y * y;
    const double e = fma(-x, t, 1.0);
    const double h = y * e;               // h and p are independent: dependency depth 4, not 5
    const double p = fma(0.375, e, 0.5);
    return fma(h, p, y);
}

// ---------------------------------------------------------------------------------------------
// theta in [0, pi] from BOTH c = cos(theta) and s = sin(theta) >= 0 (the angle code has both: the dot product and the
// norm of the cross product).  A few FP32 / integer instructions pick a tabulated angle theta_e within 0.011 rad (see
// PM_ACOS_K above: index by cos in the middle of the range, by sin near 0 and pi, where acos is ill-conditioned in cos); then
// sin(theta - theta_e) = s cos(theta_e) - c sin(theta_e) is exact to FP64 rounding and well conditioned for every theta, and
// theta - theta_e = asin(.) needs a 5-term odd series since |.| < 0.02.
// 10 FP64 ops + ~15 FP32/integer ops (idle pipes here) instead of ~50 FP64 ops for libm acos; no acosf (25 instructions) either.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double pm_acos_cs(double c, double s, const double2 *__restrict__ tab) {
    const float cf = __double2float_rn(c);
    const float a = fminf(fabsf(cf), 1.0f);
    float sf;   // ~ sin(theta) for the index only
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sf) : "f"(fmaf(-a, a, 1.0f)));
    const int ia = __float2int_rn(cf * (float)PM_ACOS_K) + PM_ACOS_NA;
    const int ib = __float2int_rn(sf * (float)PM_ACOS_K) + (cf < 0.0f ? 3 * PM_ACOS_NA + 2 : 2 * PM_ACOS_NA + 1);
    int e = a <= 0.70710678f ? ia : ib;
    e = min(max(e, 0), PM_ACOS_ENTRIES - 1);   // (NaN inputs must not index outside the table)
    const double2 t = tab[2 * e], th = tab[2 * e + 1];
    const double sd = fma(s, t.x, -(c * t.y));
    const double z = sd * sd;
    double p = fma(z, 35.0 / 1152.0, 15.0 / 336.0);
    p = fma(z, p, 3.0 / 40.0);
    p = fma(z, p, 1.0 / 6.0);
    const double delta = fma(sd * z, p, sd);
    return th.x + delta;
}

#include <cmath>
#include <vector>
// Host side: the table as 2 * (PM_ACOS_N + 1) doubles (pm_create, pm_lls_create).
static inline std::vector<double> pm_acos_table_host() {
    std::vector<double> tab(2 * (size_t)(PM_ACOS_N + 1), 0.0);
    for (int e = 0; e < PM_ACOS_ENTRIES; ++e) {
        double th;
        if (e <= 2 * PM_ACOS_NA)
            th = std::acos((double)(e - PM_ACOS_NA) / PM_ACOS_K);
        else if (e <= 3 * PM_ACOS_NA + 1)
            th = std::asin((double)(e - (2 * PM_ACOS_NA + 1)) / PM_ACOS_K);
        else
            th = PM_PI - std::asin((double)(e - (3 * PM_ACOS_NA + 2)) / PM_ACOS_K);
        tab[4 * e] = std::cos(th);
        tab[4 * e + 1] = std::sin(th);
        tab[4 * e + 2] = th;
    }
    return tab;
}

struct pm_d3 {
    double x, y, z;
};
__device__ __forceinline__ pm_d3 pm_sub(const pm_d3 a, const pm_d3 b) { return pm_d3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double pm_dot(const pm_d3 a, const pm_d3 b) { return fma(a.z, b.z, fma(a.y, b.y, a.x * b.x)); }
__device__ __forceinline__ pm_d3 pm_cross(const pm_d3 a, const pm_d3 b) {
    return pm_d3{fma(a.y, b.z, -(a.z * b.y)), fma(a.z, b.x, -(a.x * b.z)), fma(a.x, b.y, -(a.y * b.x))};
}
// a += s * v
__device__ __forceinline__ void pm_axpy(pm_d3 &a, const double s, const pm_d3 v) {
    a.x = fma(s, v.x, a.x);
    a.y = fma(s, v.y, a.y);
    a.z = fma(s, v.z, a.z);
}

// Tile accessors.  Tile = doubles [3N][CPW] (row = CPW doubles).  The topology program stores atoms as BYTE offsets of
// their first row (atom * 3 * CPW * 8), and `base` already points at this lane's column: address = base + offset.
template <int CPW>
__device__ __forceinline__ pm_d3 pm_ldo(const double *base, int boff) {
    const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(base) + boff);
    return pm_d3{p[0], p[CPW], p[2 * CPW]};
}
template <int CPW>
__device__ __forceinline__ void pm_addo(double *base, int boff, const pm_d3 v) {
    double *p = reinterpret_cast<double *>(reinterpret_cast<char *>(base) + boff);
    p[0] += v.x;
    p[CPW] += v.y;
    p[2 * CPW] += v.z;
}
template <int CPW>
__device__ __forceinline__ void pm_subo(double *base, int boff, const pm_d3 v) {
    double *p = reinterpret_cast<double *>(reinterpret_cast<char *>(base) + boff);
    p[0] -= v.x;
    p[CPW] -= v.y;
    p[2 * CPW] -= v.z;
}
// Index-based variants: tile[(3*atom + c) * CPW + column]
template <int CPW>
__device__ __forceinline__ pm_d3 pm_ld3(const double *tile, int atom, int col) {
    const double *p = tile + (3 * atom) * CPW + col;
    return pm_d3{p[0], p[CPW], p[2 * CPW]};
}
template <int CPW>
__device__ __forceinline__ void pm_add3(double *tile, int atom, int col, const pm_d3 v) {
    double *p = tile + (3 * atom) * CPW + col;
    p[0] += v.x;
    p[CPW] += v.y;
    p[2 * CPW] += v.z;
}
template <int CPW>
__device__ __forceinline__ void pm_sub3(double *tile, int atom, int col, const pm_d3 v) {
    double *p = tile + (3 * atom) * CPW + col;
    p[0] -= v.x;
    p[CPW] -= v.y;
    p[2 * CPW] -= v.z;
}
