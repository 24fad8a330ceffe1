// shell_jk.cuh — definitions shared by the shell-quartet engine (shell4.cuh) and its host side: per-angular-momentum component tables,
// the primitive shell-pair record, the device view of the shell-pair data, the lane-parallel Boys function, the NB-interleaved shared
// memory vector type and the reproducible (integer) accumulation of J/K.
//
// The reference evaluates every Cartesian-COMPONENT quartet from scratch (Boys function, R table and a six-deep Hermite loop per
// component quartet, TUNA/tuna_integrals/tuna_integral.pyx:1142-1253, driver :1312-1342) and has no direct mode.  In the engine one
// cooperative GROUP of G lanes owns a batch of SHELL quartets (AB|CD): the Boys values, the z Coulomb-Hermite table and the x/y
// convolution table are formed once and shared by all ncart(A) ncart(B) ncart(C) ncart(D) components; the integrals are folded into
// J/K blocks and flushed once per shell quartet.  The N^4 tensor is never materialised.
//
// Math (all centres on the z axis; unnormalised Cartesian Gaussians, shell-level contraction coefficients; the per-component norms f
// are folded into the density and the result outside the kernel):
//   (ab|cd) = cc_AB cc_CD 2 pi^(5/2) / (p q sqrt(p+q))
//             * sum_{m,m'} XY[ax+bx][cx+dx][m] XY[ay+by][cy+dy][m'] S[(az,bz),(cz,dz)][m+m']
//   XY[n12][n34][m] = (2m-1)!! (-1)^n34 sum_{t+tau=2m} E^{n12}_t(p) E^{n34}_tau(q)       (one-centre x/y Hermite)
//   S[bz,gz][n]     = sum_{v,phi} Ez_AB[az][bz][v] (-1)^phi Ez_CD[cz][dz][phi] R^n_{v+phi}
//   R^n_w           = sum_k a(w,k) PQz^(w-2k) B[n+w-k],  B[m] = (-2 rho)^m F_m(rho PQz^2)
#pragma once
#include "eri_core.cuh"

#if !defined(__CUDACC__)
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
#endif

namespace tuna {

constexpr int SH_LMAX = 5;
constexpr int SH_NCMAX = 21;     // (L+1)(L+2)/2 at L = 5

// Per-angular-momentum component tables (canonical order of tuna_molecule.py:622: (i, j, L-i-j), i desc, j desc).
struct ShellTab {
    int nc[SH_LMAX + 1];
    int lx[SH_LMAX + 1][SH_NCMAX], ly[SH_LMAX + 1][SH_NCMAX], lz[SH_LMAX + 1][SH_NCMAX];
    int pg[SH_LMAX + 1][SH_NCMAX];            // x/y parity code (lx&1)*2 + (ly&1)
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
    }
}

// Primitive shell-pair record (doubles): [0] p, [1] Pz, [2] c_a c_b, [3] unused,
//   [4 ...]      Ez[(az (Lb+1) + bz) (Lab+1) + v]      two-centre z Hermite coefficients, v <= az + bz
//   [4 + nEz ..] Ex[n (Lab/2+1) + t'] = E^n_t(p), t = (n&1) + 2 t'   one-centre x/y coefficients, n <= Lab
constexpr int SP_HDR = 4;
TUNA_HD int sp_ez_size(int La, int Lb) { return (La + 1) * (Lb + 1) * (La + Lb + 1); }
TUNA_HD int sp_ex_size(int La, int Lb) { return (La + Lb + 1) * ((La + Lb) / 2 + 1); }
TUNA_HD int sp_rec_size(int La, int Lb) { return SP_HDR + sp_ez_size(La, Lb) + sp_ex_size(La, Lb); }

struct ShellData {
    const int* pairA; const int* pairB;     // shell ids (A carries La >= Lb)
    const long long* pair_rec;              // offset of the first primitive record (doubles)
    const double* rec;
    const double* pairQ;                    // Schwarz factor of the shell pair (max over components, normalised integrals)
    const int* sh_ao;                       // [shell * SH_NCMAX + component] -> AO (Cartesian basis function) index
    const double* boys;
    const double* herm;
    long long fix_lo;                       // reproducible accumulation: Jf / Kf are 64-bit integer arrays, the low words live fix_lo elements behind
                                            // the high words (fixed_split below)
};

// Reproducible J/K accumulation.  Floating-point atomics make the sum depend on the order in which CTAs of six streams arrive; integer
// addition does not.  A contribution v is split exactly into  v = h * 2^-20 + r,  h = rint(v * 2^20),  and r is rounded to
// l = rint(r * 2^60) (|l| <= 2^39); h and l are added with 64-bit integer atomics into a high and a low word.  The final value
// hi * 2^-20 + lo * 2^-60 is independent of the summation order; its error is at most (contributions per element) * 2^-61, i.e. about
// 2e-14 for the 4e4 shell-pair contributions an element of K receives at nbf 800; |J|, |K| < 2^42 is required (no overflow of hi).
constexpr double FIX_HI = 1048576.0;                    // 2^20
constexpr double FIX_LO = 1152921504606846976.0;        // 2^60
TUNA_HD void fixed_split(double v, long long& h, long long& l) {
    const double hd = rint(v * FIX_HI);
    h = (long long)hd;
    l = (long long)rint(fma(-hd, 1.0 / FIX_HI, v) * FIX_LO);
}
TUNA_HD double fixed_value(long long h, long long l) { return (double)h * (1.0 / FIX_HI) + (double)l * (1.0 / FIX_LO); }

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
    double odd = 1.0;
    for (int k = 0; k < m; ++k) { f = (odd * f - e) * inv2T; odd += 2.0; }
    return f;
}

// The NB quartets a group works on are INTERLEAVED in shared memory: entry x of array X of quartet q lives at
// sm[(J.oX + x) * NB + q], so one 16-byte shared load (NB = 2) brings the operand of both quartets and every table-driven
// address is computed once per entry instead of once per quartet.
template <int NB>
struct alignas(NB >= 2 ? 16 : 8) QVec { double v[NB]; };
template <int NB>
TUNA_HD QVec<NB> qld(const double* p) { return *reinterpret_cast<const QVec<NB>*>(p); }
template <int NB>
TUNA_HD void qst(double* p, const QVec<NB>& x) { *reinterpret_cast<QVec<NB>*>(p) = x; }

}  // namespace tuna
