/*
 * oracle/eri_oracle.c — CPU restatement of the reference's two-electron-integral algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (tuna_b200/) may import, link or call this
 * file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and only as the checker.  Parity status: PINNED — tests/test_oracle.py checks it against the
 * reference's own compiled engine (oracle/_ref, built by oracle/build_ref.sh from the unmodified
 * /root/reference/TUNA/tuna_integrals/tuna_integral.pyx), against the known-answer vectors of
 * SURVEY.md section 8(c) and against the committed fixtures under tests/golden/.
 *
 * What is restated (file = TUNA/tuna_integrals/tuna_integral.pyx of h-brough/TUNA v0.12.0):
 *   oracle_normalize      <- Basis.normalize                         :174-210
 *   hermite_row           <- fill_hermite_table_iter_eri             :961-1036
 *   pair table build      <- build_primitive_pair_eri / build_ao_pair_eri   :1050-1128
 *   boys_table            <- boys / fill_boys_table                  :1490-1505, :1540-1572
 *   primitive quartet     <- fill_pow_table, fill_Rz_linear_table, primitive_pair_eri :1582-1651, :1142-1221
 *   oracle_eri_fill       <- calculate_electron_repulsion_integrals  :1267-1355
 *   oracle_eri_single     <- calculate_electron_repulsion_integral   :1376-1414
 *
 * Third-party arithmetic absent from /root/reference: the reference obtains the top-order Boys
 * function from SciPy's cython_special.hyp1f1 (pyx:10, :1505; scipy>=1.15, 1.18.1 in this image).
 * This file restates the published definition F_m(T) = 1F1(m+1/2; m+3/2; -T)/(2m+1) with the Kummer
 * series (all-positive terms) for T < 35 and erf + upward recursion above, both good to ~1e-15
 * relative; SURVEY.md section 7 measures SciPy's own values as good to ~1e-14 (8.9e-14 outlier),
 * which is why ERI parity is stated at 1e-12 Eh absolute / 1e-13 relative rather than bit-exact.
 *
 * Geometry: atoms and diatomics on the z axis (TUNA/tuna_util.py:845-878), so the x and y Hermite
 * expansions are one-centre (R = 0) and the Coulomb-Hermite recursion is one-dimensional in z.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_HERMITE 20   /* hermite_x/y/z[20], pyx:42-44: L <= 5 per function, so t <= 10 */

typedef struct {
    double coef;        /* N_a N_b c_a c_b                      pyx:1070 */
    double p;           /* a + b                                 pyx:1071 */
    double Pz;          /* (a Az + b Bz)/p                       pyx:1072 */
    double ex[ORACLE_MAX_HERMITE];
    double ey[ORACLE_MAX_HERMITE];
    double ez[ORACLE_MAX_HERMITE];
} prim_pair_t;

typedef struct {
    long i, j;
    int lx, ly, lz;     /* sums over the two functions           pyx:1106-1108 */
    int nprim;
    prim_pair_t *pp;
} ao_pair_t;

static double dfact(int n) {            /* n!! with n <= 0 -> 1, pyx:245-270 */
    double r = 1.0;
    while (n > 1) { r *= n; n -= 2; }
    return r;
}

/* Basis.normalize, pyx:174-210.  coefs is overwritten with the contraction-normalised values,
 * norm receives the primitive norms; the effective primitive coefficient is norm[k]*coefs[k]. */
void oracle_normalize(int l, int m, int n, long nprim, const double *exps, double *coefs, double *norm) {
    const double PI = 3.141592653589793238462643383279;
    int L = l + m + n;
    double dl = dfact(2 * l - 1), dm = dfact(2 * m - 1), dn = dfact(2 * n - 1);
    for (long i = 0; i < nprim; i++)
        norm[i] = sqrt(pow(2, 2 * L + 1.5) * pow(exps[i], L + 1.5) / dl / dm / dn / pow(PI, 1.5));
    double prefactor = pow(PI, 1.5) * dl * dm * dn / pow(2.0, L);
    double N = 0.0;
    for (long i = 0; i < nprim; i++)
        for (long j = 0; j < nprim; j++)
            N += norm[i] * norm[j] * coefs[i] * coefs[j] / pow(exps[i] + exps[j], L + 1.5);
    N = 1 / sqrt(prefactor * N);
    for (long i = 0; i < nprim; i++) coefs[i] *= N;
}

/* Hermite expansion coefficients E_t^{l1 l2}, t = 0..l1+l2, built bottom-up exactly in the
 * reference's order (first index raised along j == 0, second index otherwise), pyx:961-1036. */
static void hermite_row(int l1, int l2, double R, double a, double b, double *out, int parity_only) {
    int nl2 = l2 + 1, nt = l1 + l2 + 1, stride = nt + 1;
    double p = a + b, mu = a * b / p, half_inv_p = 1.0 / (2.0 * p);
    double shift1 = -mu * R / a, shift2 = mu * R / b;
    double *E = (double *)calloc((size_t)(l1 + 1) * nl2 * stride, sizeof(double));
    E[0] = exp(-mu * R * R);
    for (int i = 0; i <= l1; i++)
        for (int j = 0; j <= l2; j++) {
            if (i == 0 && j == 0) continue;
            double *cur = E + (size_t)(i * nl2 + j) * stride;
            const double *prev = (j == 0) ? E + (size_t)((i - 1) * nl2 + j) * stride
                                          : E + (size_t)(i * nl2 + j - 1) * stride;
            double shift = (j == 0) ? shift1 : shift2;
            for (int t = 0; t <= i + j; t++) {
                cur[t] = shift * prev[t] + (t + 1) * prev[t + 1];
                if (t > 0) cur[t] += half_inv_p * prev[t - 1];
            }
        }
    const double *row = E + (size_t)(l1 * nl2 + l2) * stride;
    for (int t = 0; t < ORACLE_MAX_HERMITE; t++) out[t] = 0.0;
    if (parity_only) {
        for (int t = (l1 + l2) & 1; t < nt; t += 2) out[t] = row[t];
    } else {
        for (int t = 0; t < nt; t++) out[t] = row[t];
    }
    free(E);
}

/* Top-order Boys function; stands in for scipy hyp1f1(m+1/2, m+3/2, -T)/(2m+1), pyx:1490-1505. */
double oracle_boys(int m, double T) {
    if (T < 35.0) {
        double term = 1.0 / (2.0 * m + 1.0), sum = term;
        for (int k = 1; k < 400; k++) {
            term *= 2.0 * T / (2.0 * m + 2.0 * k + 1.0);
            sum += term;
            if (term < 1e-17 * sum) break;
        }
        return exp(-T) * sum;
    }
    double e = exp(-T);
    double F = 0.5 * sqrt(3.141592653589793238462643383279 / T) * erf(sqrt(T));
    for (int k = 0; k < m; k++) F = ((2.0 * k + 1.0) * F - e) / (2.0 * T);
    return F;
}

/* fill_boys_table, pyx:1540-1572: T == 0 exact branch, otherwise top order then downward recursion. */
static void boys_table(int M, double T, double *F) {
    if (T == 0.0) {
        for (int m = 0; m <= M; m++) F[m] = 1.0 / (2.0 * m + 1.0);
        return;
    }
    F[M] = oracle_boys(M, T);
    double e = exp(-T), twoT = 2.0 * T;
    for (int m = M; m > 0; m--) F[m - 1] = (twoT * F[m] + e) / (2.0 * m - 1.0);
}

static double odd_dfact_even(int n_even) { return n_even <= 0 ? 1.0 : dfact(n_even - 1); }   /* pyx:914-947 */

/* primitive_pair_eri, pyx:1142-1221 (same loop nest and accumulation order). */
static double primitive_quartet(const ao_pair_t *A, const prim_pair_t *a, const ao_pair_t *B, const prim_pair_t *b) {
    double p = a->p, q = b->p, pq = p + q, rho = p * q / pq, PQz = a->Pz - b->Pz;
    int Vmax = A->lz + B->lz;
    int Nmax = A->lx + A->ly + A->lz + B->lx + B->ly + B->lz;
    int stride = Nmax + 1;
    double F[64], pw[64], R[1024];
    boys_table(Nmax, rho * PQz * PQz, F);
    pw[0] = 1.0;                                                   /* fill_pow_table, pyx:1582-1602 */
    for (int n = 1; n <= Nmax; n++) pw[n] = pw[n - 1] * (-2.0 * rho);
    for (int n = 0; n <= Nmax; n++) R[n] = pw[n] * F[n];           /* fill_Rz_linear_table, pyx:1612-1651 */
    for (int v = 1; v <= Vmax; v++)
        for (int n = Nmax - v; n >= 0; n--) {
            R[v * stride + n] = PQz * R[(v - 1) * stride + n + 1];
            if (v > 1) R[v * stride + n] += (v - 1) * R[(v - 2) * stride + n + 1];
        }
    double sum = 0.0;
    for (int t = A->lx & 1; t <= A->lx; t += 2)
        for (int tau = B->lx & 1; tau <= B->lx; tau += 2) {
            double xf = a->ex[t] * b->ex[tau] * odd_dfact_even(t + tau);
            for (int u = A->ly & 1; u <= A->ly; u += 2)
                for (int nu = B->ly & 1; nu <= B->ly; nu += 2) {
                    double xyf = xf * a->ey[u] * b->ey[nu] * odd_dfact_even(u + nu);
                    int nxy = ((t + tau) >> 1) + ((u + nu) >> 1);
                    for (int v = 0; v <= A->lz; v++) {
                        double ezv = a->ez[v];
                        if (ezv == 0.0) continue;
                        for (int phi = 0; phi <= B->lz; phi++) {
                            double ezp = b->ez[phi];
                            if (ezp == 0.0) continue;
                            double sign = ((tau + nu + phi) & 1) ? -1.0 : 1.0;
                            sum += xyf * ezv * ezp * sign * R[(v + phi) * stride + nxy];
                        }
                    }
                }
        }
    double prefactor = 34.986836655249725 / (p * q * sqrt(pq));    /* 2 pi^(5/2), pyx:1219 */
    return a->coef * b->coef * prefactor * sum;
}

static double contracted(const ao_pair_t *A, const ao_pair_t *B) {   /* pyx:1235-1253 */
    double s = 0.0;
    for (int i = 0; i < A->nprim; i++)
        for (int j = 0; j < B->nprim; j++) s += primitive_quartet(A, &A->pp[i], B, &B->pp[j]);
    return s;
}

typedef struct {
    long ncart;
    const double *oz;      /* origin z per function */
    const int *lmn;        /* [ncart][3] */
    const long *nprim;
    const long *off;
    const double *exps, *coefs, *norms;
} basis_view_t;

static int build_pair(ao_pair_t *P, long i, long j, const basis_view_t *B) {   /* pyx:1091-1128, :1050-1077 */
    const int *si = B->lmn + 3 * i, *sj = B->lmn + 3 * j;
    P->i = i; P->j = j;
    P->lx = si[0] + sj[0]; P->ly = si[1] + sj[1]; P->lz = si[2] + sj[2];
    P->nprim = (int)(B->nprim[i] * B->nprim[j]);
    P->pp = (prim_pair_t *)malloc(sizeof(prim_pair_t) * (size_t)P->nprim);
    if (!P->pp) return 1;
    int k = 0;
    for (long a = 0; a < B->nprim[i]; a++)
        for (long b = 0; b < B->nprim[j]; b++, k++) {
            long ia = B->off[i] + a, jb = B->off[j] + b;
            double ea = B->exps[ia], eb = B->exps[jb];
            prim_pair_t *pp = &P->pp[k];
            pp->coef = B->norms[ia] * B->norms[jb] * B->coefs[ia] * B->coefs[jb];
            pp->p = ea + eb;
            pp->Pz = (ea * B->oz[i] + eb * B->oz[j]) / pp->p;
            hermite_row(si[0], sj[0], 0.0, ea, eb, pp->ex, 1);
            hermite_row(si[1], sj[1], 0.0, ea, eb, pp->ey, 1);
            hermite_row(si[2], sj[2], B->oz[i] - B->oz[j], ea, eb, pp->ez, 0);
        }
    return 0;
}

/* calculate_electron_repulsion_integrals, pyx:1267-1355.  out is the dense C-order ncart^4 tensor.
 * Returns 0, or 1 on allocation failure (the reference raises MemoryError, pyx:1120, :1290). */
int oracle_eri_fill(long ncart, const double *origins_z, const int *lmn, const long *nprim, const long *offsets,
                    const double *exps, const double *coefs, const double *norms, double *out, int nthreads) {
    basis_view_t B = {ncart, origins_z, lmn, nprim, offsets, exps, coefs, norms};
    long npair = ncart * (ncart + 1) / 2;
    ao_pair_t *pairs = (ao_pair_t *)calloc((size_t)npair, sizeof(ao_pair_t));
    if (!pairs) return 1;
    int fail = 0;
    long idx = 0;
    for (long i = 0; i < ncart; i++)
        for (long j = 0; j <= i; j++) fail |= build_pair(&pairs[idx++], i, j, &B);
    if (!fail) {
        long n = ncart, n2 = n * n, n3 = n2 * n;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic) num_threads(nthreads > 0 ? nthreads : 1)
#endif
        for (long a = 0; a < npair; a++) {
            long i = pairs[a].i, j = pairs[a].j;
            for (long b = 0; b <= a; b++) {
                long k = pairs[b].i, l = pairs[b].j;
                double v = 0.0;
                if (!(((pairs[a].lx + pairs[b].lx) & 1) || ((pairs[a].ly + pairs[b].ly) & 1)))   /* pyx:1324-1327 */
                    v = contracted(&pairs[a], &pairs[b]);
                out[i * n3 + j * n2 + k * n + l] = v; out[k * n3 + l * n2 + i * n + j] = v;      /* pyx:1335-1342 */
                out[j * n3 + i * n2 + l * n + k] = v; out[l * n3 + k * n2 + j * n + i] = v;
                out[j * n3 + i * n2 + k * n + l] = v; out[l * n3 + k * n2 + i * n + j] = v;
                out[i * n3 + j * n2 + l * n + k] = v; out[k * n3 + l * n2 + j * n + i] = v;
            }
        }
    }
    for (long a = 0; a < npair; a++) free(pairs[a].pp);
    free(pairs);
    return fail;
}

/* calculate_electron_repulsion_integral, pyx:1376-1414: one quartet (f0 f1 | f2 f3) of the same flattened basis. */
double oracle_eri_single(const double *origins_z, const int *lmn, const long *nprim, const long *offsets,
                         const double *exps, const double *coefs, const double *norms, long f0, long f1, long f2, long f3) {
    basis_view_t B = {0, origins_z, lmn, nprim, offsets, exps, coefs, norms};
    ao_pair_t A, C;
    double v = 0.0;
    if (build_pair(&A, f0, f1, &B)) return NAN;
    if (build_pair(&C, f2, f3, &B)) { free(A.pp); return NAN; }
    if (!(((A.lx + C.lx) & 1) || ((A.ly + C.ly) & 1))) v = contracted(&A, &C);
    free(A.pp); free(C.pp);
    return v;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
