// shell_jk.cuh — shell-quartet engine for direct J/K on a z-axis diatomic (the headline kernel).
//
// The reference evaluates every Cartesian-COMPONENT quartet from scratch (Boys function, R table and a
// six-deep Hermite loop per component quartet, TUNA/tuna_integrals/tuna_integral.pyx:1142-1253, driver
// :1312-1342) and has no direct mode.  Here one cooperative GROUP of G lanes owns one SHELL quartet
// (AB|CD): the Boys values, the z Coulomb-Hermite table and the x/y convolution table are formed once
// and shared by all ncart(A) ncart(B) ncart(C) ncart(D) components, and the integrals are folded into
// shared-memory J/K blocks that are flushed with one atomic per block entry per shell quartet (instead
// of six global atomics per component quartet).  The N^4 tensor is never materialised.
//
// Math (all centres on the z axis; unnormalised Cartesian Gaussians, shell-level contraction coefficients;
// the per-component norms f are folded into the density and the result on the host side of the kernel):
//   (ab|cd) = cc_AB cc_CD 2 pi^(5/2) / (p q sqrt(p+q))
//             * sum_{m,m'} XY[ax+bx][cx+dx][m] XY[ay+by][cy+dy][m'] S[(az,bz),(cz,dz)][m+m']
//   XY[n12][n34][m] = (2m-1)!! (-1)^n34 sum_{t+tau=2m} E^{n12}_t(p) E^{n34}_tau(q)       (one-centre x/y Hermite)
//   S[bz,gz][n]     = sum_{v,phi} Ez_AB[az][bz][v] (-1)^phi Ez_CD[cz][dz][phi] R^n_{v+phi}
//   R^n_w           = sum_k a(w,k) PQz^(w-2k) B[n+w-k],  B[m] = (-2 rho)^m F_m(rho PQz^2)
//
// The body is written against a Policy (lane id, group size, group barrier, atomic add) so that the same
// source is the sm_100a kernel (DevPolicy<G>) and, with G = 1, a CPU unit-test build (tests/host_emul).
#pragma once
#include "eri_core.cuh"

namespace tuna {

constexpr int SH_LMAX = 5;
constexpr int SH_NCMAX = 21;     // (L+1)(L+2)/2 at L = 5

// Per-angular-momentum component tables (canonical order of tuna_molecule.py:622: (i, j, L-i-j), i desc, j desc).
struct ShellTab {
    int nc[SH_LMAX + 1];
    int lx[SH_LMAX + 1][SH_NCMAX], ly[SH_LMAX + 1][SH_NCMAX], lz[SH_LMAX + 1][SH_NCMAX];
    int pg[SH_LMAX + 1][SH_NCMAX];            // x/y parity code (lx&1)*2 + (ly&1)
    int goff[SH_LMAX + 1][5];                 // components sorted by parity code: group offsets
    int glist[SH_LMAX + 1][SH_NCMAX];         // component ids in group order
    int gslot[SH_LMAX + 1][SH_NCMAX];         // position of a component inside its group
    int gmax[SH_LMAX + 1];                    // largest group
    int zoff[SH_LMAX + 1][SH_LMAX + 2];       // components sorted by lz: offsets
    int zlist[SH_LMAX + 1][SH_NCMAX];
};

inline void build_shell_tab(ShellTab& T) {
    for (int L = 0; L <= SH_LMAX; ++L) {
        int c = 0;
        for (int i = L; i >= 0; --i)
            for (int j = L - i; j >= 0; --j, ++c) {
                T.lx[L][c] = i; T.ly[L][c] = j; T.lz[L][c] = L - i - j;
                T.pg[L][c] = (i & 1) * 2 + (j & 1);
            }
        T.nc[L] = c;
        int pos = 0;
        T.gmax[L] = 0;
        for (int g = 0; g < 4; ++g) {
            T.goff[L][g] = pos;
            for (int k = 0; k < c; ++k)
                if (T.pg[L][k] == g) { T.glist[L][pos] = k; T.gslot[L][k] = pos - T.goff[L][g]; ++pos; }
            if (pos - T.goff[L][g] > T.gmax[L]) T.gmax[L] = pos - T.goff[L][g];
        }
        T.goff[L][4] = pos;
        pos = 0;
        for (int z = 0; z <= L; ++z) {
            T.zoff[L][z] = pos;
            for (int k = 0; k < c; ++k)
                if (T.lz[L][k] == z) T.zlist[L][pos++] = k;
        }
        T.zoff[L][L + 1] = pos;
    }
}

// Primitive shell-pair record (doubles): [0] p, [1] Pz, [2] c_a c_b, [3] unused,
//   [4 ...]      Ez[(az (Lb+1) + bz) (Lab+1) + v]      two-centre z Hermite coefficients, v <= az + bz
//   [4 + nEz ..] Ex[n (Lab/2+1) + t'] = E^n_t(p), t = (n&1) + 2 t'   one-centre x/y coefficients, n <= Lab
constexpr int SP_HDR = 4;
TUNA_HD int sp_ez_size(int La, int Lb) { return (La + 1) * (Lb + 1) * (La + Lb + 1); }
TUNA_HD int sp_ex_size(int La, int Lb) { return (La + Lb + 1) * ((La + Lb) / 2 + 1); }
TUNA_HD int sp_rec_size(int La, int Lb) { return SP_HDR + sp_ez_size(La, Lb) + sp_ex_size(La, Lb); }

// One launch = one (bra pair class, ket pair class) job.
struct ShellJob {
    int La, Lb, Lc, Ld;
    int nppAB, nppCD;               // primitive pairs per shell pair (uniform inside a class)
    const int* bra_list;            // pair ids of the bra class, Schwarz-descending
    const int* ket_list;
    const long long* item_prefix;   // [nbra + 1]: kets kept per bra (Schwarz cut, and ket_pos <= bra_pos if same class)
    int nbra, same_class;
    long long nitems;
    // shared-memory layout of one group (offsets in doubles)
    int NS, NGZ, oB, oPz, oRt, oXY, oU, oS, oIt, oKAC, oKAD, oKBC, oKBD, oJAB, oJCD, total;
};

struct ShellData {
    const int* pairA; const int* pairB;     // shell ids (A carries La >= Lb)
    const long long* pair_rec;              // offset of the first primitive record (doubles)
    const double* rec;
    const double* pairQ;                    // Schwarz factor of the shell pair (max over components, normalised integrals)
    const int* sh_ao;                       // [shell * SH_NCMAX + component] -> AO (Cartesian basis function) index
    const ShellTab* tab;
    const double* boys;
    const double* herm;
};

inline void shell_job_layout(ShellJob& J, const ShellTab& T, int nD) {
    const int Ltot = J.La + J.Lb + J.Lc + J.Ld, Lab = J.La + J.Lb, Lcd = J.Lc + J.Ld;
    J.NS = Ltot / 2 + 1;
    J.NGZ = (J.Lc + 1) * (J.Ld + 1);
    int o = 0;
    J.oB = o; o += Ltot + 1;
    J.oPz = o; o += Ltot + 1;
    J.oRt = o; o += (Ltot + 1) * J.NS;
    J.oXY = o; o += (Lab + 1) * (Lcd + 1) * J.NS;
    J.oU = o; o += (Lab + 1) * J.NGZ * J.NS;
    J.oS = o; o += J.NGZ * J.NS;
    J.oIt = o; o += (J.La + 1) * (J.Lb + 1) * T.nc[J.Lc] * T.gmax[J.Ld];
    J.oKAC = o; o += nD * T.nc[J.La] * T.nc[J.Lc];
    J.oKAD = o; o += nD * T.nc[J.La] * T.nc[J.Ld];
    J.oKBC = o; o += nD * T.nc[J.Lb] * T.nc[J.Lc];
    J.oKBD = o; o += nD * T.nc[J.Lb] * T.nc[J.Ld];
    J.oJAB = o; o += nD * T.nc[J.La] * T.nc[J.Lb];
    J.oJCD = o; o += nD * T.nc[J.Lc] * T.nc[J.Ld];
    J.total = (o + 1) & ~1;
}

// F_m(T) for ONE order (lane-parallel Boys): Taylor about the table row for T < BOYS_TMAX, upward recursion above.
TUNA_HD double boys_single(const double* __restrict__ tab, int m, double T) {
    if (T < (double)BOYS_TMAX) {
        int i = (int)(T * BOYS_INV_STEP + 0.5);
        double d = (double)i * (1.0 / BOYS_INV_STEP) - T;
        const double* row = tab + (size_t)i * BOYS_COLS + m;
        double s = row[BOYS_TAYLOR];
        s = fma(s, d * (1.0 / 8.0), row[7]);
        s = fma(s, d * (1.0 / 7.0), row[6]);
        s = fma(s, d * (1.0 / 6.0), row[5]);
        s = fma(s, d * (1.0 / 5.0), row[4]);
        s = fma(s, d * (1.0 / 4.0), row[3]);
        s = fma(s, d * (1.0 / 3.0), row[2]);
        s = fma(s, d * (1.0 / 2.0), row[1]);
        return fma(s, d, row[0]);
    }
    double inv2T = 0.5 / T, e = exp(-T);
    double f = 0.88622692545275801365 * sqrt(1.0 / T);
    for (int k = 0; k < m; ++k) f = ((double)(2 * k + 1) * f - e) * inv2T;
    return f;
}

#define TUNA_LANES(i, n) for (int i = Pol::lane(); i < (n); i += Pol::G)

// One shell quartet (pair ids AB, CD; degeneracy weight w) folded into the global accumulators Jf, Kf
// (nD matrices of ncart x ncart each) for densities Pf.  `sm` is this group's private shared-memory slice.
// `active` = false groups only take part in the barriers.
template <class Pol>
TUNA_HD void shell_quartet(const ShellJob& J, const ShellData& D, bool active, int AB, int CD, double w, double* __restrict__ sm,
                           int nD, const double* __restrict__ Pf, double* Jf, double* Kf, int ncart) {
    const ShellTab& T = *D.tab;
    const int La = J.La, Lb = J.Lb, Lc = J.Lc, Ld = J.Ld;
    const int Lab = La + Lb, Lcd = Lc + Ld, Ltot = Lab + Lcd, NS = J.NS, NGZ = J.NGZ;
    const int ncA = T.nc[La], ncB = T.nc[Lb], ncC = T.nc[Lc], ncD = T.nc[Ld], gmaxD = T.gmax[Ld];
    const int NTA = Lab / 2 + 1, NTC = Lcd / 2 + 1;
    const size_t nn = (size_t)ncart * ncart;
    double* B = sm + J.oB; double* pzt = sm + J.oPz; double* Rt = sm + J.oRt; double* XY = sm + J.oXY;
    double* U = sm + J.oU; double* S = sm + J.oS; double* It = sm + J.oIt;
    double* KAC = sm + J.oKAC; double* KAD = sm + J.oKAD; double* KBC = sm + J.oKBC; double* KBD = sm + J.oKBD;
    double* JAB = sm + J.oJAB; double* JCD = sm + J.oJCD;

    int shA = 0, shB = 0, shC = 0, shD = 0;
    const double* recA = nullptr; const double* recC = nullptr;
    if (active) {
        shA = D.pairA[AB]; shB = D.pairB[AB]; shC = D.pairA[CD]; shD = D.pairB[CD];
        recA = D.rec + D.pair_rec[AB]; recC = D.rec + D.pair_rec[CD];
        TUNA_LANES(x, J.total - J.oKAC) KAC[x] = 0.0;      // all six accumulator blocks are contiguous
    }
    const int* aoA = D.sh_ao + shA * SH_NCMAX; const int* aoB = D.sh_ao + shB * SH_NCMAX;
    const int* aoC = D.sh_ao + shC * SH_NCMAX; const int* aoD = D.sh_ao + shD * SH_NCMAX;
    const int recAsz = sp_rec_size(La, Lb), recCsz = sp_rec_size(Lc, Ld);
    const int oExA = SP_HDR + sp_ez_size(La, Lb), oExC = SP_HDR + sp_ez_size(Lc, Ld);

    for (int ia = 0; ia < J.nppAB; ++ia)
        for (int ic = 0; ic < J.nppCD; ++ic) {
            const double* rA = recA + (size_t)ia * recAsz;
            const double* rC = recC + (size_t)ic * recCsz;
            double pref = 0.0;
            // ---- phase 0: Boys values scaled by (-2 rho)^m, powers of PQz -------------------------------------
            if (active) {
                const double p = rA[0], q = rC[0], pq = p + q, rho = p * q / pq, PQz = rA[1] - rC[1];
                const double Targ = rho * PQz * PQz;
                pref = w * rA[2] * rC[2] * 34.986836655249725 / (p * q * sqrt(pq));
                TUNA_LANES(m, Ltot + 1) {
                    double f = boys_single(D.boys, m, Targ), s = 1.0, z = 1.0;
                    for (int k = 0; k < m; ++k) { s *= -2.0 * rho; z *= PQz; }
                    B[m] = f * s;
                    pzt[m] = z;
                }
            }
            Pol::sync();
            // ---- phase 1: R^n_w (closed form) and the x/y convolution table --------------------------------------
            if (active) {
                TUNA_LANES(x, (Ltot + 1) * NS) {
                    const int wv = x / NS, n = x % NS;
                    if (2 * n + wv > Ltot) continue;
                    double r = 0.0;
                    for (int k = 0; 2 * k <= wv; ++k) r = fma(D.herm[wv * HERM_STRIDE + k] * pzt[wv - 2 * k], B[n + wv - k], r);
                    Rt[x] = r;
                }
                const double* ExA = rA + oExA; const double* ExC = rC + oExC;
                TUNA_LANES(x, (Lab + 1) * (Lcd + 1) * NS) {
                    const int m = x % NS, n34 = (x / NS) % (Lcd + 1), n12 = x / (NS * (Lcd + 1));
                    const int px = n12 & 1;
                    double v = 0.0;
                    if (((n12 ^ n34) & 1) == 0 && m >= px && 2 * m <= n12 + n34) {
                        // t = px + 2 t', tau = 2m - t = px + 2 tau'
                        const int tlo = (2 * m - n34 > px) ? 2 * m - n34 : px;
                        const int thi = (2 * m - px < n12) ? 2 * m - px : n12;
                        for (int t = tlo; t <= thi; t += 2) v = fma(ExA[n12 * NTA + (t >> 1)], ExC[n34 * NTC + ((2 * m - t) >> 1)], v);
                        v *= odd_dfact(m);
                        if (n34 & 1) v = -v;
                    }
                    XY[x] = v;
                }
            }
            Pol::sync();
            // ---- phase 2: U[v][gz][n] = sum_phi (-1)^phi Ez_CD[gz][phi] R^n_{v+phi} -------------------------------
            if (active) {
                const double* EzC = rC + SP_HDR;
                TUNA_LANES(x, (Lab + 1) * NGZ * NS) {
                    const int n = x % NS, gz = (x / NS) % NGZ, v = x / (NS * NGZ);
                    const int lz34 = gz / (Ld + 1) + gz % (Ld + 1);
                    if (2 * n + v + lz34 > Ltot) continue;
                    const double* e = EzC + gz * (Lcd + 1);
                    double u = 0.0;
                    for (int phi = 0; phi <= lz34; ++phi) {
                        const double t = e[phi] * Rt[(v + phi) * NS + n];
                        u = (phi & 1) ? u - t : u + t;
                    }
                    U[x] = u;
                }
            }
            Pol::sync();
            // ---- loop over bra z-combinations --------------------------------------------------------------------
            for (int az = 0; az <= La; ++az)
                for (int bz = 0; bz <= Lb; ++bz) {
                    const int lz12 = az + bz;
                    const int nA = La - az + 1, nB = Lb - bz + 1;          // components of A with lz = az, of B with lz = bz
                    const int* zA = T.zlist[La] + T.zoff[La][az];
                    const int* zB = T.zlist[Lb] + T.zoff[Lb][bz];
                    // phase 3: S[gz][n] = sum_v Ez_AB[az][bz][v] U[v][gz][n]
                    if (active) {
                        const double* e = rA + SP_HDR + (az * (Lb + 1) + bz) * (Lab + 1);
                        TUNA_LANES(x, NGZ * NS) {
                            const int n = x % NS, gz = x / NS;
                            const int lz34 = gz / (Ld + 1) + gz % (Ld + 1);
                            if (2 * n + lz12 + lz34 > Ltot) continue;
                            double s = 0.0;
                            for (int v = 0; v <= lz12; ++v) s = fma(e[v], U[(v * NGZ + gz) * NS + n], s);
                            S[x] = s;
                        }
                    }
                    Pol::sync();
                    // phase 4: the integrals of this slice, It[a'][b'][c][slot of d in its parity group]
                    if (active) {
                        TUNA_LANES(x, nA * nB * ncC * gmaxD) {
                            const int slot = x % gmaxD, c = (x / gmaxD) % ncC, bp = (x / (gmaxD * ncC)) % nB, ap = x / (gmaxD * ncC * nB);
                            const int a = zA[ap], b = zB[bp];
                            const int g = T.pg[La][a] ^ T.pg[Lb][b] ^ T.pg[Lc][c];
                            if (slot >= T.goff[Ld][g + 1] - T.goff[Ld][g]) continue;
                            const int d = T.glist[Ld][T.goff[Ld][g] + slot];
                            const int nx12 = T.lx[La][a] + T.lx[Lb][b], nx34 = T.lx[Lc][c] + T.lx[Ld][d];
                            const int ny12 = T.ly[La][a] + T.ly[Lb][b], ny34 = T.ly[Lc][c] + T.ly[Ld][d];
                            const int gz = T.lz[Lc][c] * (Ld + 1) + T.lz[Ld][d];
                            const double* xr = XY + (nx12 * (Lcd + 1) + nx34) * NS;
                            const double* yr = XY + (ny12 * (Lcd + 1) + ny34) * NS;
                            const double* sr = S + gz * NS;
                            double val = 0.0;
                            for (int m = nx12 & 1; 2 * m <= nx12 + nx34; ++m) {
                                double t = 0.0;
                                for (int mp = ny12 & 1; 2 * mp <= ny12 + ny34; ++mp) t = fma(yr[mp], sr[m + mp], t);
                                val = fma(xr[m], t, val);
                            }
                            It[((ap * (Lb + 1) + bp) * ncC + c) * gmaxD + slot] = pref * val;
                        }
                    }
                    Pol::sync();
                    // phase 5: digestion.  Each output entry is owned by one lane; the six blocks are disjoint.
                    if (active) {
                        for (int dn = 0; dn < nD; ++dn) {
                            const double* P = Pf + dn * nn;
                            // KAC[a][c] += sum_{b,d} I P[d][b]      KBC[b][c] += sum_{a,d} I P[d][a]
                            TUNA_LANES(x, (nA + nB) * ncC) {
                                const int c = x % ncC, r = x / ncC;
                                const bool isA = r < nA;
                                const int ap0 = isA ? r : 0, ap1 = isA ? r + 1 : nA, bp0 = isA ? 0 : r - nA, bp1 = isA ? nB : r - nA + 1;
                                double s = 0.0;
                                for (int ap = ap0; ap < ap1; ++ap)
                                    for (int bp = bp0; bp < bp1; ++bp) {
                                        const int a = zA[ap], b = zB[bp];
                                        const int g = T.pg[La][a] ^ T.pg[Lb][b] ^ T.pg[Lc][c];
                                        const int g0 = T.goff[Ld][g], gn = T.goff[Ld][g + 1] - g0;
                                        const double* it = It + ((ap * (Lb + 1) + bp) * ncC + c) * gmaxD;
                                        const int col = isA ? aoB[b] : aoA[a];
                                        for (int sl = 0; sl < gn; ++sl) s = fma(it[sl], P[(size_t)aoD[T.glist[Ld][g0 + sl]] * ncart + col], s);
                                    }
                                if (isA) KAC[(dn * ncA + zA[r]) * ncC + c] += s;
                                else KBC[(dn * ncB + zB[r - nA]) * ncC + c] += s;
                            }
                            // KAD[a][d] += sum_{b,c} I P[c][b]      KBD[b][d] += sum_{a,c} I P[c][a]
                            TUNA_LANES(x, (nA + nB) * ncD) {
                                const int d = x % ncD, r = x / ncD;
                                const bool isA = r < nA;
                                const int ap0 = isA ? r : 0, ap1 = isA ? r + 1 : nA, bp0 = isA ? 0 : r - nA, bp1 = isA ? nB : r - nA + 1;
                                const int sl = T.gslot[Ld][d];
                                double s = 0.0;
                                for (int ap = ap0; ap < ap1; ++ap)
                                    for (int bp = bp0; bp < bp1; ++bp) {
                                        const int a = zA[ap], b = zB[bp];
                                        const int g = T.pg[La][a] ^ T.pg[Lb][b] ^ T.pg[Ld][d];       // parity group of c
                                        const int g0 = T.goff[Lc][g], g1 = T.goff[Lc][g + 1];
                                        const double* it = It + ((ap * (Lb + 1) + bp) * ncC) * gmaxD + sl;
                                        const int col = isA ? aoB[b] : aoA[a];
                                        for (int k = g0; k < g1; ++k) {
                                            const int c = T.glist[Lc][k];
                                            s = fma(it[c * gmaxD], P[(size_t)aoC[c] * ncart + col], s);
                                        }
                                    }
                                if (isA) KAD[(dn * ncA + zA[r]) * ncD + d] += s;
                                else KBD[(dn * ncB + zB[r - nA]) * ncD + d] += s;
                            }
                            // JCD[c][d] += sum_{a,b} I (P[a][b] + P[b][a])
                            TUNA_LANES(x, ncC * ncD) {
                                const int d = x % ncD, c = x / ncD;
                                const int gcd = T.pg[Lc][c] ^ T.pg[Ld][d], sl = T.gslot[Ld][d];
                                double s = 0.0;
                                for (int ap = 0; ap < nA; ++ap)
                                    for (int bp = 0; bp < nB; ++bp) {
                                        const int a = zA[ap], b = zB[bp];
                                        if ((T.pg[La][a] ^ T.pg[Lb][b]) != gcd) continue;
                                        const double pab = P[(size_t)aoA[a] * ncart + aoB[b]] + P[(size_t)aoB[b] * ncart + aoA[a]];
                                        s = fma(It[((ap * (Lb + 1) + bp) * ncC + c) * gmaxD + sl], pab, s);
                                    }
                                JCD[(dn * ncC + c) * ncD + d] += s;
                            }
                            // JAB[a][b] += sum_{c,d} I (P[c][d] + P[d][c])
                            TUNA_LANES(x, nA * nB) {
                                const int bp = x % nB, ap = x / nB;
                                const int a = zA[ap], b = zB[bp];
                                const int gab = T.pg[La][a] ^ T.pg[Lb][b];
                                const double* it = It + ((ap * (Lb + 1) + bp) * ncC) * gmaxD;
                                double s = 0.0;
                                for (int c = 0; c < ncC; ++c) {
                                    const int g = gab ^ T.pg[Lc][c];
                                    const int g0 = T.goff[Ld][g], gn = T.goff[Ld][g + 1] - g0;
                                    for (int sl = 0; sl < gn; ++sl) {
                                        const int d = T.glist[Ld][g0 + sl];
                                        const double pcd = P[(size_t)aoC[c] * ncart + aoD[d]] + P[(size_t)aoD[d] * ncart + aoC[c]];
                                        s = fma(it[c * gmaxD + sl], pcd, s);
                                    }
                                }
                                JAB[(dn * ncA + a) * ncB + b] += s;
                            }
                        }
                    }
                    Pol::sync();
                }
        }
    // ---- flush the shell blocks: one atomic per block entry per shell quartet ------------------------------------
    if (active) {
        for (int dn = 0; dn < nD; ++dn) {
            double* Jd = Jf + dn * nn;
            double* Kd = Kf + dn * nn;
            TUNA_LANES(x, ncA * ncC) Pol::atomic_add(Kd + (size_t)aoA[x / ncC] * ncart + aoC[x % ncC], KAC[dn * ncA * ncC + x]);
            TUNA_LANES(x, ncA * ncD) Pol::atomic_add(Kd + (size_t)aoA[x / ncD] * ncart + aoD[x % ncD], KAD[dn * ncA * ncD + x]);
            TUNA_LANES(x, ncB * ncC) Pol::atomic_add(Kd + (size_t)aoB[x / ncC] * ncart + aoC[x % ncC], KBC[dn * ncB * ncC + x]);
            TUNA_LANES(x, ncB * ncD) Pol::atomic_add(Kd + (size_t)aoB[x / ncD] * ncart + aoD[x % ncD], KBD[dn * ncB * ncD + x]);
            TUNA_LANES(x, ncA * ncB) Pol::atomic_add(Jd + (size_t)aoA[x / ncB] * ncart + aoB[x % ncB], JAB[dn * ncA * ncB + x]);
            TUNA_LANES(x, ncC * ncD) Pol::atomic_add(Jd + (size_t)aoC[x / ncD] * ncart + aoD[x % ncD], JCD[dn * ncC * ncD + x]);
        }
    }
    Pol::sync();
}

// item index -> (bra position, ket position) through the per-bra prefix of kept kets
TUNA_HD void shell_item_decode(const ShellJob& J, long long item, int& ib, int& ik) {
    int lo = 0, hi = J.nbra;        // invariant: prefix[lo] <= item < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (J.item_prefix[mid] <= item) lo = mid; else hi = mid;
    }
    ib = lo;
    ik = (int)(item - J.item_prefix[lo]);
}

}  // namespace tuna
