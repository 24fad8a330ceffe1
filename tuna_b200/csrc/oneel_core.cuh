// oneel_core.cuh — one-electron integrals of a pair of contracted Cartesian Gaussians on the z axis.
//
// Replaces calculate_contracted_local_integrals (TUNA/tuna_integrals/tuna_integral.pyx:446-625: overlap, kinetic energy, dipole and
// diagonal quadrupole moments), calculate_contracted_nuclear_integral (:779-912) and the recursive hermite_coeff (:1428-1489, called
// ~20 times per primitive pair, each an exponential recursion) with ONE iterative pass per Cartesian direction: the Hermite rows
// E_t^{l1,j} are built upward in j with a two-row buffer, and the three rows the formulas need (j = l2 - 2, l2, l2 + 2) are kept.
// The nuclear attraction uses the same Boys table and Coulomb-Hermite recursion as the two-electron kernels (eri_core.cuh).
// Everything is TUNA_HD: compiled for the device by tuna_b200.cu and for the host by tests/host_emul (parity on the CPU).
#pragma once
#include "eri_core.cuh"

namespace tuna {

constexpr int OE_NT = 16;          // Hermite index range: t <= l1 + l2 + 2 <= 12
constexpr int OE_LMAX = 5;
constexpr int OE_NR = 2 * OE_LMAX + 1;      // orders of the Coulomb-Hermite table: n, v <= 10

// E_t^{l1,l2}(t = 0..l1+l2) into E[], and the t = 0 coefficients of the rows (l1, l2 + 2) and (l1, l2 - 2).
// Same recursion as pyx:1428-1489 with x_PA = -b/p R, x_PB = a/p R.
TUNA_HD void oe_hermite_rows(int l1, int l2, double R, double a, double b, double* E, double& e0_plus2, double& e0_minus2) {
    const double p = a + b, mu = a * b / p, h = 0.5 / p, xpa = -b / p * R, xpb = a / p * R;
    double r0[OE_NT], r1[OE_NT];
    for (int t = 0; t < OE_NT; ++t) { r0[t] = 0.0; r1[t] = 0.0; }
    double* prev = r0;
    double* cur = r1;
    prev[0] = exp(-mu * R * R);
    for (int i = 1; i <= l1; ++i) {                         // (i, 0) from (i - 1, 0)
        for (int t = 0; t <= i; ++t) {
            double v = xpa * prev[t] + (double)(t + 1) * prev[t + 1];
            if (t > 0) v += h * prev[t - 1];
            cur[t] = v;
        }
        double* s = prev; prev = cur; cur = s;
    }
    e0_minus2 = 0.0;
    for (int j = 0;; ++j) {                                  // prev holds row (l1, j)
        if (j == l2 - 2) e0_minus2 = prev[0];
        if (j == l2)
            for (int t = 0; t <= l1 + l2; ++t) E[t] = prev[t];
        if (j == l2 + 2) { e0_plus2 = prev[0]; break; }
        for (int t = 0; t <= l1 + j + 1; ++t) {
            double v = xpb * prev[t] + (double)(t + 1) * prev[t + 1];
            if (t > 0) v += h * prev[t - 1];
            cur[t] = v;
        }
        cur[l1 + j + 2] = 0.0;
        double* s = prev; prev = cur; cur = s;
    }
}

struct OneElPair { double s, t, v, d[3], q[3]; };

// All one-electron integrals of the AO pair (A, B).  lmnA / lmnB: Cartesian powers; zA / zB: centres; expsX / ceffX: primitives with
// ceff = norm * coef (pyx:504-508); atoms on the z axis with charges Z; origin[3]: dipole / quadrupole origin; boys: device Boys table.
TUNA_HD OneElPair one_electron_pair(const int* lmnA, double zA, int nA, const double* expsA, const double* ceffA, const int* lmnB, double zB, int nB,
                                    const double* expsB, const double* ceffB, int natoms, const double* atom_z, const double* atom_charge,
                                    const double* origin, const double* __restrict__ boys) {
    OneElPair o;
    o.s = o.t = o.v = 0.0;
    for (int c = 0; c < 3; ++c) { o.d[c] = 0.0; o.q[c] = 0.0; }
    const double R[3] = {0.0, 0.0, zA - zB};
    const int Vmax = lmnA[2] + lmnB[2], Nmax = lmnA[0] + lmnB[0] + lmnA[1] + lmnB[1] + Vmax;
    for (int i = 0; i < nA; ++i)
        for (int j = 0; j < nB; ++j) {
            const double a = expsA[i], b = expsB[j], p = a + b;
            const double pref = ceffA[i] * ceffB[j] * 5.5683279968317078452848179821188357 / (p * sqrt(p));      // pi^(3/2), pyx:14
            double E[3][OE_NT], S[3], T1[3], D1[3], Q1[3];
            const double Pc[3] = {0.0 - origin[0], 0.0 - origin[1], (a * zA + b * zB) / p - origin[2]};
            for (int c = 0; c < 3; ++c) {
                const int l1 = lmnA[c], l2 = lmnB[c];
                double ep2, em2;
                oe_hermite_rows(l1, l2, R[c], a, b, E[c], ep2, em2);
                const double e1 = (l1 + l2 >= 1) ? E[c][1] : 0.0, e2 = (l1 + l2 >= 2) ? E[c][2] : 0.0;
                S[c] = E[c][0];
                T1[c] = (double)(2 * l2 + 1) * b * S[c] - 2.0 * b * b * ep2 - 0.5 * (double)(l2 * (l2 - 1)) * em2;     // pyx:541-549
                D1[c] = e1 + Pc[c] * S[c];                                                                              // pyx:557-559
                Q1[c] = 2.0 * e2 + 2.0 * Pc[c] * e1 + (Pc[c] * Pc[c] + 1.0 / (2.0 * p)) * S[c];                           // pyx:563-565
            }
            o.s += pref * S[0] * S[1] * S[2];
            o.t += pref * (T1[0] * S[1] * S[2] + S[0] * T1[1] * S[2] + S[0] * S[1] * T1[2]);
            o.d[0] += pref * D1[0] * S[1] * S[2]; o.d[1] += pref * S[0] * D1[1] * S[2]; o.d[2] += pref * S[0] * S[1] * D1[2];
            o.q[0] += pref * Q1[0] * S[1] * S[2]; o.q[1] += pref * S[0] * Q1[1] * S[2]; o.q[2] += pref * S[0] * S[1] * Q1[2];
            // nuclear attraction: sum over nuclei of -Z <A| 1/|r - C| |B>, C on the z axis (pyx:779-912)
            const double Pz = (a * zA + b * zB) / p;
            double vsum = 0.0;
            for (int at = 0; at < natoms; ++at) {
                const double PCz = Pz - atom_z[at];
                double F[2 * OE_NR], Rz[OE_NR * OE_NR];
                boys_fill(boys, Nmax, p * PCz * PCz, F);
                double pw = 1.0;
                for (int n = 0; n <= Nmax; ++n) { Rz[n] = pw * F[n]; pw *= -2.0 * p; }              // R_0^n = (-2p)^n F_n
                for (int v = 1; v <= Vmax; ++v)
                    for (int n = Nmax - v; n >= 0; --n) {
                        double r = PCz * Rz[(v - 1) * OE_NR + n + 1];
                        if (v > 1) r += (double)(v - 1) * Rz[(v - 2) * OE_NR + n + 1];
                        Rz[v * OE_NR + n] = r;
                    }
                double prim = 0.0;
                for (int t = 0; t <= lmnA[0] + lmnB[0]; t += 2) {
                    const double ex = E[0][t] * odd_dfact(t / 2);                                   // (t - 1)!!
                    for (int u = 0; u <= lmnA[1] + lmnB[1]; u += 2) {
                        const double exy = ex * E[1][u] * odd_dfact(u / 2);
                        const double* rz = Rz + (t + u) / 2;
                        double acc = 0.0;
                        for (int v = 0; v <= Vmax; ++v) acc += E[2][v] * rz[v * OE_NR];
                        prim += exy * acc;
                    }
                }
                vsum -= atom_charge[at] * prim;
            }
            o.v += ceffA[i] * ceffB[j] * vsum * 6.283185307179586476925286766559 / p;
        }
    return o;
}

// Overlap only (cross-basis overlap, pyx:626-778).
TUNA_HD double overlap_pair(const int* lmnA, double zA, int nA, const double* expsA, const double* ceffA, const int* lmnB, double zB, int nB,
                            const double* expsB, const double* ceffB) {
    double s = 0.0;
    for (int i = 0; i < nA; ++i)
        for (int j = 0; j < nB; ++j) {
            const double a = expsA[i], b = expsB[j], p = a + b;
            double prod = ceffA[i] * ceffB[j] * 5.5683279968317078452848179821188357 / (p * sqrt(p));
            for (int c = 0; c < 3; ++c) {
                double E[OE_NT], ep2, em2;
                oe_hermite_rows(lmnA[c], lmnB[c], c == 2 ? zA - zB : 0.0, a, b, E, ep2, em2);
                prod *= E[0];
            }
            s += prod;
        }
    return s;
}

}  // namespace tuna
