/*
 * oracle/oneel_oracle.c — CPU restatement of the reference's one-electron integrals.
 *
 * TEST INFRASTRUCTURE ONLY (same rules as eri_oracle.c: tests/, smoke() and bench.py's CPU legs only, as the checker).
 * Parity status: PINNED — tests/test_one_electron.py compares it with the reference's own compiled engine
 * (oracle/_ref: calculate_one_electron_integrals / calculate_cross_basis_overlap_matrix of the unmodified
 * TUNA/tuna_integrals/tuna_integral.pyx) on every committed configuration.
 *
 * What is restated (file = TUNA/tuna_integrals/tuna_integral.pyx of h-brough/TUNA v0.12.0), loop for loop:
 *   hermite_coeff                          <- hermite_coeff (recursive)                       :1428-1489
 *   local_integrals                        <- calculate_contracted_local_integrals            :446-625   (S, T, dipole, quadrupole diagonals)
 *   nuclear_integral                       <- calculate_contracted_nuclear_integral           :779-912
 *   oracle_one_electron                    <- calculate_one_electron_integrals                :282-445
 *   oracle_cross_overlap                   <- calculate_cross_basis_overlap_matrix            :626-778
 * The Boys function comes from eri_oracle.c (oracle_boys: the Kummer-series stand-in for SciPy's hyp1f1, pyx:1505).
 * Geometry: all centres on the z axis (the reference's nuclear integral is only valid there, pyx:783).
 */
#include <math.h>
#include <stdlib.h>

double oracle_boys(int m, double T);       /* eri_oracle.c */

static const double PI = 3.141592653589793238462643383279;     /* pyx:13 */
static const double PI32 = 5.5683279968317078452848179821188357; /* pyx:14 */

static double double_fact(int n) {          /* pyx:245-270: n!! with n <= 0 -> 1 */
    double r = 1.0;
    for (; n > 1; n -= 2) r *= n;
    return r;
}

static double hermite_coeff(int l1, int l2, int t, double R, double a, double b) {     /* pyx:1428-1489 */
    const double p = a + b, u = a * b / p, pre = 1.0 / (2.0 * p);
    if (t < 0 || t > l1 + l2) return 0.0;
    if (l1 == 0 && l2 == 0 && t == 0) return exp(-u * R * R);
    if (l2 == 0)
        return pre * hermite_coeff(l1 - 1, l2, t - 1, R, a, b) - (u * R / a) * hermite_coeff(l1 - 1, l2, t, R, a, b) +
               (t + 1) * hermite_coeff(l1 - 1, l2, t + 1, R, a, b);
    return pre * hermite_coeff(l1, l2 - 1, t - 1, R, a, b) + (u * R / b) * hermite_coeff(l1, l2 - 1, t, R, a, b) +
           (t + 1) * hermite_coeff(l1, l2 - 1, t + 1, R, a, b);
}

typedef struct {
    double z;
    int l, m, n;
    long nprim;
    const double *exps, *ceff;      /* ceff = norm[k] * coefs[k] (pyx:504-508) */
} bf_t;

/* out[8] = s, t, dx, dy, dz, qxx, qyy, qzz                                              pyx:446-625 */
static void local_integrals(const bf_t *A, const bf_t *B, const double *origin, double *out) {
    const double dx = 0.0, dy = 0.0, dz = A->z - B->z;
    for (int k = 0; k < 8; k++) out[k] = 0.0;
    for (long i = 0; i < A->nprim; i++)
        for (long j = 0; j < B->nprim; j++) {
            const double a = A->exps[i], b = B->exps[j], p = a + b;
            const double pref = A->ceff[i] * B->ceff[j] * PI32 / (p * sqrt(p));
            const double Sx = hermite_coeff(A->l, B->l, 0, dx, a, b), Sy = hermite_coeff(A->m, B->m, 0, dy, a, b), Sz = hermite_coeff(A->n, B->n, 0, dz, a, b);
            const double Ex1 = hermite_coeff(A->l, B->l, 1, dx, a, b), Ey1 = hermite_coeff(A->m, B->m, 1, dy, a, b), Ez1 = hermite_coeff(A->n, B->n, 1, dz, a, b);
            const double Ex2 = hermite_coeff(A->l, B->l, 2, dx, a, b), Ey2 = hermite_coeff(A->m, B->m, 2, dy, a, b), Ez2 = hermite_coeff(A->n, B->n, 2, dz, a, b);
            const double Ax = (2 * B->l + 1) * b, Ay = (2 * B->m + 1) * b, Az = (2 * B->n + 1) * b;
            const double Bx = -0.5 * B->l * (B->l - 1), By = -0.5 * B->m * (B->m - 1), Bz = -0.5 * B->n * (B->n - 1);
            const double Tx = Ax * Sx - 2.0 * b * b * hermite_coeff(A->l, B->l + 2, 0, dx, a, b) + Bx * hermite_coeff(A->l, B->l - 2, 0, dx, a, b);
            const double Ty = Ay * Sy - 2.0 * b * b * hermite_coeff(A->m, B->m + 2, 0, dy, a, b) + By * hermite_coeff(A->m, B->m - 2, 0, dy, a, b);
            const double Tz = Az * Sz - 2.0 * b * b * hermite_coeff(A->n, B->n + 2, 0, dz, a, b) + Bz * hermite_coeff(A->n, B->n - 2, 0, dz, a, b);
            const double Px = 0.0 - origin[0], Py = 0.0 - origin[1], Pz = (a * A->z + b * B->z) / p - origin[2];
            const double Dx = Ex1 + Px * Sx, Dy = Ey1 + Py * Sy, Dz = Ez1 + Pz * Sz;
            const double Qx = 2.0 * Ex2 + 2.0 * Px * Ex1 + (Px * Px + 1.0 / (2.0 * p)) * Sx;
            const double Qy = 2.0 * Ey2 + 2.0 * Py * Ey1 + (Py * Py + 1.0 / (2.0 * p)) * Sy;
            const double Qz = 2.0 * Ez2 + 2.0 * Pz * Ez1 + (Pz * Pz + 1.0 / (2.0 * p)) * Sz;
            out[0] += pref * Sx * Sy * Sz;
            out[1] += pref * (Tx * Sy * Sz + Sx * Ty * Sz + Sx * Sy * Tz);
            out[2] += pref * Dx * Sy * Sz;
            out[3] += pref * Sx * Dy * Sz;
            out[4] += pref * Sx * Sy * Dz;
            out[5] += pref * Qx * Sy * Sz;
            out[6] += pref * Sx * Qy * Sz;
            out[7] += pref * Sx * Sy * Qz;
        }
}

/* <1| 1/|r - C| |2> for a nucleus on the z axis                                         pyx:779-912 */
static double nuclear_integral(const bf_t *A, const bf_t *B, double zc) {
    const double Rz12 = A->z - B->z;
    const int Vmax = A->n + B->n, Nmax = A->l + B->l + A->m + B->m + A->n + B->n, stride = Nmax + 1;
    double *F = malloc((size_t)(Nmax + 1) * sizeof(double)), *pw = malloc((size_t)(Nmax + 1) * sizeof(double));
    double *Rz = malloc((size_t)(Vmax + 1) * (Nmax + 1) * sizeof(double));
    double integral = 0.0;
    for (long i = 0; i < A->nprim; i++)
        for (long j = 0; j < B->nprim; j++) {
            const double a = A->exps[i], b = B->exps[j], p = a + b;
            const double PCz = (a * A->z + b * B->z) / p - zc, T = p * PCz * PCz;
            if (T == 0.0) {                                                    /* fill_boys_table, pyx:1540-1572 */
                for (int m = 0; m <= Nmax; m++) F[m] = 1.0 / (2.0 * m + 1.0);
            } else {
                F[Nmax] = oracle_boys(Nmax, T);
                const double e = exp(-T), twoT = 2.0 * T;
                for (int m = Nmax; m > 0; m--) F[m - 1] = (twoT * F[m] + e) / (2.0 * m - 1.0);
            }
            pw[0] = 1.0;                                                       /* fill_pow_table, pyx:1582-1602 */
            for (int n = 1; n <= Nmax; n++) pw[n] = pw[n - 1] * (-2.0 * p);
            for (int n = 0; n <= Nmax; n++) Rz[n] = pw[n] * F[n];              /* fill_Rz_linear_table, pyx:1612-1651 */
            for (int v = 1; v <= Vmax; v++)
                for (int n = Nmax - v; n >= 0; n--) {
                    Rz[v * stride + n] = PCz * Rz[(v - 1) * stride + n + 1];
                    if (v > 1) Rz[v * stride + n] += (v - 1) * Rz[(v - 2) * stride + n + 1];
                }
            double prim = 0.0;
            for (int t = 0; t <= A->l + B->l; t += 2) {
                const double Ex = hermite_coeff(A->l, B->l, t, 0.0, a, b) * double_fact(t - 1);
                for (int u = 0; u <= A->m + B->m; u += 2) {
                    const double Ey = hermite_coeff(A->m, B->m, u, 0.0, a, b) * double_fact(u - 1);
                    for (int v = 0; v <= A->n + B->n; v++) {
                        const double Ez = hermite_coeff(A->n, B->n, v, Rz12, a, b);
                        prim += Ex * Ey * Ez * Rz[v * stride + (t + u) / 2];
                    }
                }
            }
            integral += A->ceff[i] * B->ceff[j] * prim * 2.0 * PI / p;
        }
    free(F); free(pw); free(Rz);
    return integral;
}

static bf_t view(long i, const double *oz, const int *lmn, const long *nprim, const long *off, const double *exps, const double *ceff) {
    bf_t b;
    b.z = oz[i]; b.l = lmn[3 * i]; b.m = lmn[3 * i + 1]; b.n = lmn[3 * i + 2];
    b.nprim = nprim[i]; b.exps = exps + off[i]; b.ceff = ceff + off[i];
    return b;
}

/* calculate_one_electron_integrals, pyx:282-445.  S, T, V: [n][n]; D, Q: [3][n][n]. */
int oracle_one_electron(long n, const double *oz, const int *lmn, const long *nprim, const long *off, const double *exps, const double *ceff,
                        long natoms, const double *atom_z, const double *atom_charge, const double *origin, double *S, double *T, double *V,
                        double *D, double *Q, int nthreads) {
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 1)
    for (long i = 0; i < n; i++)
        for (long j = 0; j <= i; j++) {
            const bf_t A = view(i, oz, lmn, nprim, off, exps, ceff), B = view(j, oz, lmn, nprim, off, exps, ceff);
            double o[8];
            local_integrals(&A, &B, origin, o);
            double v = 0.0;
            for (long a = 0; a < natoms; a++) v = v - nuclear_integral(&A, &B, atom_z[a]) * atom_charge[a];
            const long ij = i * n + j, ji = j * n + i, nn = n * n;
            S[ij] = S[ji] = o[0];
            T[ij] = T[ji] = o[1];
            V[ij] = V[ji] = v;
            for (int c = 0; c < 3; c++) {
                D[c * nn + ij] = D[c * nn + ji] = o[2 + c];
                Q[c * nn + ij] = Q[c * nn + ji] = o[5 + c];
            }
        }
    return 0;
}

/* calculate_cross_basis_overlap_matrix, pyx:626-778.  S12: [n1][n2]. */
int oracle_cross_overlap(long n1, const double *oz1, const int *lmn1, const long *nprim1, const long *off1, const double *exps1, const double *ceff1,
                         long n2, const double *oz2, const int *lmn2, const long *nprim2, const long *off2, const double *exps2, const double *ceff2,
                         double *S12) {
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n1; i++)
        for (long j = 0; j < n2; j++) {
            const bf_t A = view(i, oz1, lmn1, nprim1, off1, exps1, ceff1), B = view(j, oz2, lmn2, nprim2, off2, exps2, ceff2);
            double s = 0.0;
            for (long k = 0; k < A.nprim; k++)
                for (long l = 0; l < B.nprim; l++) {
                    const double a = A.exps[k], b = B.exps[l], p = a + b;
                    const double pref = A.ceff[k] * B.ceff[l] * PI32 / (p * sqrt(p));
                    s += pref * hermite_coeff(A.l, B.l, 0, 0.0, a, b) * hermite_coeff(A.m, B.m, 0, 0.0, a, b) * hermite_coeff(A.n, B.n, 0, A.z - B.z, a, b);
                }
            S12[i * n2 + j] = s;
        }
    return 0;
}
