// eri_core.cuh — FP64 math shared by every ERI kernel: Boys function, pair-record layout and the
// primitive-quartet integral for a z-axis diatomic.
//
// Everything here is `TUNA_HD` (= __host__ __device__ under nvcc, nothing under g++) so that the
// identical source is (a) inlined into the sm_100a kernels of tuna_b200.cu and (b) compiled by g++
// into tests/host_emul (a CPU unit-test harness for this development container, which has no GPU;
// it is never loaded by the package and is not a fallback).
//
// Reference being replaced (TUNA/tuna_integrals/tuna_integral.pyx of h-brough/TUNA):
//   boys_fill        <- boys + fill_boys_table            :1490-1505, :1540-1572
//   eri_prim_quartet <- primitive_pair_eri                :1142-1221 (with fill_pow_table :1582, fill_Rz_linear_table :1612)
// The arithmetic is re-derived, not transcribed: because every centre is on the z axis the x/y Hermite
// indices only enter through n = (t+tau)/2 + (u+nu)/2, so the reference's six nested loops factor into
//   integral = sum_m B[m] * (G * C)[m],   B[m] = (-2 rho)^m F_m(T)
// with G = the x/y convolution (geometry independent) and C = the z Hermite polynomial in PQz.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define TUNA_HD __host__ __device__ __forceinline__
#else
#define TUNA_HD inline
#endif

namespace tuna {

// ---------------------------------------------------------------------------------------------
// Boys function table: F_m(T_i), T_i = i / BOYS_INV_STEP, i = 0..BOYS_ROWS-1, m = 0..BOYS_COLS-1.
// Top order from an 8th-order Taylor expansion about the nearest grid point (dF_m/dT = -F_{m+1}),
// lower orders by the stable downward recursion; T >= BOYS_TMAX uses F_0 = sqrt(pi/T)/2 and the
// upward recursion, which is stable there for m <= 20.  Accuracy ~2e-16 relative (tests/test_host_emul.py).
// ---------------------------------------------------------------------------------------------
constexpr int BOYS_MMAX = 20;        // 4 * L_max, L_max = 5 ("H" shells, tuna_molecule.py:612-618)
constexpr int BOYS_TAYLOR = 8;
constexpr int BOYS_COLS = 32;        // >= BOYS_MMAX + BOYS_TAYLOR + 1
constexpr int BOYS_INV_STEP = 16;
constexpr int BOYS_TMAX = 40;
constexpr int BOYS_ROWS = BOYS_TMAX * BOYS_INV_STEP + 1;

TUNA_HD void boys_fill(const double* __restrict__ tab, int M, double T, double* F) {
    if (T < (double)BOYS_TMAX) {
        int i = (int)(T * BOYS_INV_STEP + 0.5);
        double d = (double)i * (1.0 / BOYS_INV_STEP) - T;
        const double* row = tab + (size_t)i * BOYS_COLS + M;
        double s = row[BOYS_TAYLOR];
        s = fma(s, d * (1.0 / 8.0), row[7]);
        s = fma(s, d * (1.0 / 7.0), row[6]);
        s = fma(s, d * (1.0 / 6.0), row[5]);
        s = fma(s, d * (1.0 / 5.0), row[4]);
        s = fma(s, d * (1.0 / 4.0), row[3]);
        s = fma(s, d * (1.0 / 3.0), row[2]);
        s = fma(s, d * (1.0 / 2.0), row[1]);
        s = fma(s, d, row[0]);
        F[M] = s;
        if (M > 0) {
            double e = exp(-T), t2 = 2.0 * T;
            for (int m = M; m > 0; --m) F[m - 1] = fma(t2, F[m], e) / (double)(2 * m - 1);
        }
    } else {
        double inv2T = 0.5 / T;
        double f = 0.88622692545275801365 * sqrt(1.0 / T);   // sqrt(pi)/2 / sqrt(T); erf(sqrt(T)) == 1 in FP64 here
        F[0] = f;
        if (M > 0) {
            double e = exp(-T);
            for (int m = 0; m < M; ++m) { f = ((double)(2 * m + 1) * f - e) * inv2T; F[m + 1] = f; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Primitive-pair record of one AO (Cartesian-component) pair — the "shell-pair / primitive table"
// built once per geometry (host, pairtable.hpp) and read by every quartet kernel.
//   ex[a] = E^x_t, t = (lx&1) + 2a   (one-centre: depends on lx = l1x + l2x and p only)
//   ey[a] likewise, ez[v] = E^z_v, v = 0..lz (two-centre, R = Az - Bz)
// Replaces PrimitivePairERI (pyx:35-45, 512 B) with 26 doubles = 208 B.
// ---------------------------------------------------------------------------------------------
constexpr int PP_EX = 3, PP_EY = 9, PP_EZ = 15, PP_DOUBLES = 26;   // [0]=coef [1]=p [2]=Pz
constexpr int L_PAIR_MAX = 10;                                     // lx, ly, lz sums of a pair <= 2 * 5

// (2m-1)!! for m = 0..10  (odd_double_fact_even_argument_fast, pyx:914-947)
TUNA_HD double odd_dfact(int m) {
    const double t[11] = {1.0, 1.0, 3.0, 15.0, 105.0, 945.0, 10395.0, 135135.0, 2027025.0, 34459425.0, 654729075.0};
    return t[m];
}

// Coefficient of PQz^(w-2k) B[n+w-k] in R_w^n (closed form of the recursion at pyx:1645-1651):
//   a(w,k) = w! / (k! (w-2k)! 2^k), stored by the caller as herm[w * 11 + k], w <= 20, k <= 10.
constexpr int HERM_STRIDE = 11;

struct PairClass { int lx, ly, lz; };

// One primitive quartet.  A, B: 26-double records.  ca/cb: pair classes.  Returns the primitive ERI
// INCLUDING coefficients and the 2 pi^(5/2) / (p q sqrt(p+q)) prefactor (pyx:1219-1221).
TUNA_HD double eri_prim_quartet(const double* __restrict__ A, const double* __restrict__ B, PairClass ca, PairClass cb,
                                const double* __restrict__ boys_tab, const double* __restrict__ herm) {
    const double p = A[1], q = B[1];
    const double pq = p + q, rho = p * q / pq, PQz = A[2] - B[2];
    const int nxa = ca.lx >> 1, nxb = cb.lx >> 1, nya = ca.ly >> 1, nyb = cb.ly >> 1;
    const int offx = ca.lx & 1, offy = ca.ly & 1;           // == cb parities (caller applied the parity test)
    const int Vmax = ca.lz + cb.lz;
    const int nG = nxa + nxb + nya + nyb;                     // G has nG + 1 entries; true Boys order = offx + offy + index
    const int off = offx + offy;

    // x and y convolutions, then G = X * Y
    double X[L_PAIR_MAX + 1], Y[L_PAIR_MAX + 1], G[2 * L_PAIR_MAX + 1];
    for (int s = 0; s <= nxa + nxb; ++s) X[s] = 0.0;
    for (int a = 0; a <= nxa; ++a)
        for (int b = 0; b <= nxb; ++b) X[a + b] = fma(A[PP_EX + a], B[PP_EX + b], X[a + b]);
    for (int s = 0; s <= nxa + nxb; ++s) X[s] *= odd_dfact(s + offx);
    for (int s = 0; s <= nya + nyb; ++s) Y[s] = 0.0;
    for (int a = 0; a <= nya; ++a)
        for (int b = 0; b <= nyb; ++b) Y[a + b] = fma(A[PP_EY + a], B[PP_EY + b], Y[a + b]);
    for (int s = 0; s <= nya + nyb; ++s) Y[s] *= odd_dfact(s + offy);
    for (int s = 0; s <= nG; ++s) G[s] = 0.0;
    for (int a = 0; a <= nxa + nxb; ++a)
        for (int b = 0; b <= nya + nyb; ++b) G[a + b] = fma(X[a], Y[b], G[a + b]);

    // z: Z[w] = sum_{v+phi=w} Ez12[v] (-1)^phi Ez34[phi]
    double Z[2 * L_PAIR_MAX + 1], C[2 * L_PAIR_MAX + 1], pz[2 * L_PAIR_MAX + 1];
    for (int w = 0; w <= Vmax; ++w) Z[w] = 0.0;
    for (int phi = 0; phi <= cb.lz; ++phi) {
        double e = (phi & 1) ? -B[PP_EZ + phi] : B[PP_EZ + phi];
        for (int v = 0; v <= ca.lz; ++v) Z[v + phi] = fma(A[PP_EZ + v], e, Z[v + phi]);
    }
    // C[j] = sum_w Z[w] a(w, w-j) PQz^(2j-w): S_n = sum_w Z[w] R_w^n = sum_j C[j] B[n+j]
    pz[0] = 1.0;
    for (int e = 1; e <= Vmax; ++e) pz[e] = pz[e - 1] * PQz;
    for (int j = 0; j <= Vmax; ++j) {
        double c = 0.0;
        int whi = (2 * j < Vmax) ? 2 * j : Vmax;
        for (int w = j; w <= whi; ++w) c = fma(Z[w] * herm[w * HERM_STRIDE + (w - j)], pz[2 * j - w], c);
        C[j] = c;
    }

    // Boys table up to the highest order actually used, scaled by (-2 rho)^m
    const int Mtop = off + nG + Vmax;
    double F[BOYS_MMAX + 1];
    boys_fill(boys_tab, Mtop, rho * PQz * PQz, F);
    double sum = 0.0;
    {
        const double m2rho = -2.0 * rho;
        double scale = 1.0;
        for (int m = 0; m < off; ++m) scale *= m2rho;
        // sum_m B[off + m] * W[m],  W = G * C
        for (int m = 0; m <= nG + Vmax; ++m) {
            int lo = (m > Vmax) ? m - Vmax : 0, hi = (m < nG) ? m : nG;
            double w = 0.0;
            for (int n = lo; n <= hi; ++n) w = fma(G[n], C[m - n], w);
            sum = fma(w * scale, F[off + m], sum);
            scale *= m2rho;
        }
    }
    const double sign = ((cb.lx + cb.ly) & 1) ? -1.0 : 1.0;     // (-1)^(tau+nu): tau = lx34 mod 2, nu = ly34 mod 2
    const double pref = 34.986836655249725 / (p * q * sqrt(pq));  // 2 pi^(5/2)
    return A[0] * B[0] * sign * pref * sum;
}

// Contracted AO quartet = sum over primitive pairs (pyx:1235-1253).
TUNA_HD double eri_ao_quartet(const double* __restrict__ ppA, int nA, const double* __restrict__ ppB, int nB, PairClass ca, PairClass cb,
                              const double* __restrict__ boys_tab, const double* __restrict__ herm) {
    double s = 0.0;
    for (int i = 0; i < nA; ++i)
        for (int j = 0; j < nB; ++j)
            s += eri_prim_quartet(ppA + (size_t)i * PP_DOUBLES, ppB + (size_t)j * PP_DOUBLES, ca, cb, boys_tab, herm);
    return s;
}

// The same sum restricted to primitive quartets pq = start, start + stride, ... (pq = i * nB + j): lets the lanes of a
// warp share one heavily contracted AO quartet (e.g. (ss|ss) of cc-pVTZ: 64 x 64 primitive quartets).
TUNA_HD double eri_ao_quartet_strided(const double* __restrict__ ppA, int nA, const double* __restrict__ ppB, int nB, PairClass ca, PairClass cb,
                                      const double* __restrict__ boys_tab, const double* __restrict__ herm, int start, int stride) {
    double s = 0.0;
    const int total = nA * nB;
    int i = start / nB, j = start % nB;
    for (int pq = start; pq < total; pq += stride) {
        s += eri_prim_quartet(ppA + (size_t)i * PP_DOUBLES, ppB + (size_t)j * PP_DOUBLES, ca, cb, boys_tab, herm);
        j += stride;
        while (j >= nB) { j -= nB; ++i; }
    }
    return s;
}

}  // namespace tuna
