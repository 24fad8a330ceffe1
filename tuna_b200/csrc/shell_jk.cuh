// shell_jk.cuh — shell-quartet engine for direct J/K on a z-axis diatomic (the headline kernel).
//
// The reference evaluates every Cartesian-COMPONENT quartet from scratch (Boys function, R table and a
// six-deep Hermite loop per component quartet, TUNA/tuna_integrals/tuna_integral.pyx:1142-1253, driver
// :1312-1342) and has no direct mode.  Here one cooperative GROUP of G lanes owns one SHELL quartet
// (AB|CD): the Boys values, the z Coulomb-Hermite table and the x/y convolution table are formed once
// and shared by all ncart(A) ncart(B) ncart(C) ncart(D) components; the integrals are folded into
// shared-memory J/K blocks and flushed with one atomic per block entry per shell quartet (instead of
// six global atomics per component quartet).  The N^4 tensor is never materialised.
//
// Math (all centres on the z axis; unnormalised Cartesian Gaussians, shell-level contraction coefficients;
// the per-component norms f are folded into the density and the result outside the kernel):
//   (ab|cd) = cc_AB cc_CD 2 pi^(5/2) / (p q sqrt(p+q))
//             * sum_{m,m'} XY[ax+bx][cx+dx][m] XY[ay+by][cy+dy][m'] S[(az,bz),(cz,dz)][m+m']
//   XY[n12][n34][m] = (2m-1)!! (-1)^n34 sum_{t+tau=2m} E^{n12}_t(p) E^{n34}_tau(q)       (one-centre x/y Hermite)
//   S[bz,gz][n]     = sum_{v,phi} Ez_AB[az][bz][v] (-1)^phi Ez_CD[cz][dz][phi] R^n_{v+phi}
//   R^n_w           = sum_k a(w,k) PQz^(w-2k) B[n+w-k],  B[m] = (-2 rho)^m F_m(rho PQz^2)
//
// All component-index arithmetic is done ONCE per angular class on the host (ClassTables): the integral
// assembly (phase 4) and the J/K digestion (phase 5) are table-driven loops, so the device code contains no
// per-integral divisions or parity logic.  The body is written against a Policy (lane id, group size, barrier,
// atomic add): DevPolicy<G> is the sm_100a kernel, the serial HostPolicy is the CPU unit-test build
// (tests/host_emul).
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

// Host-built, per angular class (La, Lb, Lc, Ld): everything that depends on component indices only.
//   Integrals are listed chunk by chunk (a chunk = a run of bra z-combinations (az,bz) whose S slice and integral
//   buffer fit the shared-memory budget).  Phase-4 entry of integral e (2 words):
//     w0 = xoff | yoff << 16          offsets of the XY rows of the x and y index pairs
//     w1 = soff | mx0<<16 | mx1<<20 | my0<<24 | my1<<28   S row offset inside the chunk slice and the m / m' ranges
//   Phase 5 is a CSR per chunk over the NOUT accumulators (K blocks KAC,KAD,KBC,KBD per component, J per bra / ket pair
//   function; sorted by work):
//     term = it_index_in_chunk | pstage_index << 16
//   The staged density blocks (same sizes/order: P[d][b], P[c][b], P[d][a], P[c][a], P[c][d]+P[d][c], P[a][b]+P[b][a])
//   and the output blocks are addressed through `pmap` / `omap`: row | col << 8 with row/col = shell_sel << 5 | component.
struct ClassTablesDev {
    int nchunk, nout, itmax, smax_rows;     // smax_rows: largest number of (az,bz) rows in a chunk
    const int* chunk_bz0;                   // [nchunk+1] first bra z-combination index of each chunk
    const int* chunk_e0;                    // [nchunk+1] first integral of each chunk
    const unsigned* p4;                     // [2 * nint]
    const unsigned* p5ptr;                  // [nchunk * (nout + 1)]
    const unsigned* p5term;                 // concatenated; chunk c starts at p5off[c]
    const unsigned* p5off;                  // [nchunk]
    const unsigned short* pmap;             // [nk]    staged K density entry -> (row, col)
    const unsigned short* omap;             // [nout]  accumulator -> (row, col) of its K entry, or 0xffff for a J pair-function accumulator
    int nk;                                 // number of K density entries: Pst[0..nk) are single-element gathers through pmap
    // J in pair-function space: component pairs (a,b) with equal (ax+bx, ay+by, az, bz) share one integral row, so the Coulomb
    // part is accumulated per pair function and expanded to components only at the flush.
    //   Pst[nk + i], i < njst : sum of the symmetrised density over the component pairs of pair function i
    //                           (jst_ptr/jst_list CSR; list entry = row | col << 8 like pmap)
    //   jflush[i], i < njfl   : (row | col << 8) | accumulator position << 16 : Jf[row][col] += Out[position]
    int njst, njfl;
    const unsigned* jst_ptr; const unsigned short* jst_list; const unsigned* jflush;
    // phases 1-3 as flat work lists (no per-entry index arithmetic on the device):
    //   t_rt[i]  = (w * NS + n) | w << 16 | n << 24                                   R^n_w entries with 2n + w <= Ltot
    //   t_xy[i]  = ((n12 (Lcd+1) + n34) NS + m) | n12 << 16 | n34 << 20 | m << 24     non-zero XY entries
    //   t_u[2i]  = ((v NGZ + gz) NS + n) | (v NS + n) << 16 ; t_u[2i+1] = gz (Lcd+1) | lz34 << 16
    //   t_s[2i]  = S offset in the chunk slice | (gz NS + n) << 16 ; t_s[2i+1] = (az (Lb+1) + bz)(Lab+1) | lz12 << 16
    int n_rt, n_xy, n_u;
    const unsigned* t_rt; const unsigned* t_xy; const unsigned* t_u; const unsigned* t_s;
    const int* chunk_s0;                    // [nchunk+1] first t_s entry of each chunk
    // dense-tensor fill (stored mode): the parity-allowed component quartets of every chunk,
    //   p6[2i] = It slot | a << 16 | b << 21 | c << 26,  p6[2i+1] = d   (component indices inside the four shells)
    const unsigned* p6;
    const int* chunk_f0;                    // [nchunk+1] first p6 entry of each chunk
#ifdef TUNA_SHELL_WIDE_TERMS
    // Development variant (off by default): phase-5 terms as two 32-bit BYTE offsets (It slot, staged density entry), already
    // multiplied by 8 NB for the job's batch size, in the same transposed 32-accumulator blocks (two 16-byte loads per quad):
    // no shift/mask decode in the digestion loop.  Both offsets are relative to the It buffer (the density word includes
    // (oP - oIt) 8 NB).  Chunk c starts at word 2 * p5off[c].
    const unsigned* p5w;
#endif
};

// One launch = one (bra pair class, ket pair class) job.
struct ShellJob {
    int La, Lb, Lc, Ld;
    int nppAB, nppCD;               // primitive pairs per shell pair (uniform inside a class)
    const int* bra_list;            // pair ids of the bra class, Schwarz-descending
    const int* ket_list;
    const long long* item_prefix;   // [nbra + 1]: kets kept per bra (Schwarz cut, and ket_pos <= bra_pos if same class)
    int nbra, same_class;
    int chunk;                      // consecutive items per CTA work unit / sharding unit
    long long nitems;
    int dbg_skip;                   // development aid: bit k set -> phase k is skipped (timing experiments only; 0 in production)
    int fill;                       // 1: no digestion - the integrals of every chunk are scattered into the dense tensor D.eri_out
    double uniq[6];                 // unique AO quartets per shell quartet by degeneracy case (ClassTablesHost::uniq)
    ClassTablesDev ct;
    // shared-memory layout of one group (offsets in doubles)
    int NS, NGZ, oB, oPz, oRt, oXY, oU, oS, oIt, oP, oOut, oRecA, oRecC, oAO, aostride, total;
};

struct ShellData {
    const int* pairA; const int* pairB;     // shell ids (A carries La >= Lb)
    const long long* pair_rec;              // offset of the first primitive record (doubles)
    const double* rec;
    const double* pairQ;                    // Schwarz factor of the shell pair (max over components, normalised integrals)
    const int* sh_ao;                       // [shell * SH_NCMAX + component] -> AO (Cartesian basis function) index
    const double* boys;
    const double* herm;
    double* eri_out;                        // fill mode: dense Cartesian tensor [ncart^4] (zero-initialised by the caller)
    const double* fnorm;                    // fill mode: per-component norms (the engine works with unnormalised components)
    long long fix_lo;                       // generation 4, reproducible accumulation: 0 = FP64 atomics into Jf / Kf; otherwise Jf / Kf are 64-bit
                                            // integer arrays and the low words live fix_lo elements behind the high words (see fixed_add)
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

inline void shell_job_layout(ShellJob& J, int nD) {
    const int Ltot = J.La + J.Lb + J.Lc + J.Ld, Lab = J.La + J.Lb, Lcd = J.Lc + J.Ld;
    J.NS = Ltot / 2 + 1;
    J.NGZ = (J.Lc + 1) * (J.Ld + 1);
    int o = 0;
    J.oB = o; o += Ltot + 1;
    J.oPz = o; o += Ltot + 1;
    J.oRt = o; o += (Ltot + 1) * J.NS;
    J.oXY = o; o += (Lab + 1) * (Lcd + 1) * J.NS;
    J.oU = o; o += (Lab + 1) * J.NGZ * J.NS;
    J.oS = o; o += J.ct.smax_rows * J.NGZ * J.NS;
    J.oIt = o; o += J.ct.itmax + 1;       // + the zero slot read by padding terms
    J.oP = o; o += nD * J.ct.nout;
    J.oOut = o; o += nD * J.ct.nout;
    J.oRecA = o; o += sp_rec_size(J.La, J.Lb);        // primitive shell-pair records of the current primitive quartet
    J.oRecC = o; o += sp_rec_size(J.Lc, J.Ld);
    int lmax = J.La > J.Lc ? J.La : J.Lc;            // La >= Lb, Lc >= Ld by construction
    if (J.Lb > lmax) lmax = J.Lb;
    if (J.Ld > lmax) lmax = J.Ld;
    J.aostride = (lmax + 1) * (lmax + 2) / 2;
    J.oAO = o; o += (4 * J.aostride * (int)sizeof(int) + 7) / 8;      // AO indices of the four shells (ints)
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
    double odd = 1.0;
    for (int k = 0; k < m; ++k) { f = (odd * f - e) * inv2T; odd += 2.0; }
    return f;
}

#define TUNA_LANES(i, n) for (int i = Pol::lane(); i < (n); i += Pol::G)

// The NB quartets a group works on are INTERLEAVED in shared memory: entry x of array X of quartet q lives at
// sm[(J.oX + x) * NB + q], so one 16-byte shared load (NB = 2) brings the operand of both quartets and every table-driven
// address is computed once per entry instead of once per quartet.
template <int NB>
struct alignas(NB >= 2 ? 16 : 8) QVec { double v[NB]; };
template <int NB>
TUNA_HD QVec<NB> qld(const double* p) { return *reinterpret_cast<const QVec<NB>*>(p); }
template <int NB>
TUNA_HD void qst(double* p, const QVec<NB>& x) { *reinterpret_cast<QVec<NB>*>(p) = x; }

#ifdef TUNA_SHELL_ASM_UNROLL
// Development variant (off by default) of the phase-4 inner loops: the NY x/y-convolution operands of the y pair are loaded into
// registers once per integral and the m' loop is fully unrolled, so every (m, m') term costs one shared load and NB FMAs instead of
// two loads, NB FMAs and the control of a 1-4 trip loop.  Same summation order as the generic loop (bit-identical results).
template <int NB, int NY>
TUNA_HD QVec<NB> assemble_integral(const double* xyx, const double* xyy, const double* s_row, int mx0, int mx1, int my0) {
    QVec<NB> y[NY];
#pragma unroll
    for (int k = 0; k < NY; ++k) y[k] = qld<NB>(xyy + (size_t)(my0 + k) * NB);
    QVec<NB> val;
#pragma unroll
    for (int q = 0; q < NB; ++q) val.v[q] = 0.0;
    for (int m = mx0; m <= mx1; ++m) {
        const double* sp = s_row + (size_t)(m + my0) * NB;
        QVec<NB> t;
#pragma unroll
        for (int q = 0; q < NB; ++q) t.v[q] = 0.0;
#pragma unroll
        for (int k = 0; k < NY; ++k) {
            const QVec<NB> sv = qld<NB>(sp + (size_t)k * NB);
#pragma unroll
            for (int q = 0; q < NB; ++q) t.v[q] = fma(y[k].v[q], sv.v[q], t.v[q]);
        }
        const QVec<NB> x = qld<NB>(xyx + (size_t)m * NB);
#pragma unroll
        for (int q = 0; q < NB; ++q) val.v[q] = fma(x.v[q], t.v[q], val.v[q]);
    }
    return val;
}
#endif

// NB shell quartets of the same class (pair ids AB[], CD[]; degeneracy weights w[]) processed TOGETHER by one group and folded
// into the global accumulators Jf, Kf (nD matrices of ncart x ncart each) for densities Pf.  Batching NB quartets amortises
// every table-entry decode, loop and barrier over NB independent FMA streams.  Quartets with active[q] == false run on the
// data of an active quartet with weight zero (uniform control flow, no per-quartet branches) and are not flushed; a batch
// without any active quartet only takes part in the barriers.
template <class Pol, int NB>
TUNA_HD void shell_quartets(const ShellJob& J, const ShellData& D, const bool* active, const int* ABin, const int* CDin, const double* win,
                            double* __restrict__ sm, int nD, const double* __restrict__ Pf, const double* __restrict__ Psym, double* Jf,
                            double* Kf, int ncart) {
    const ClassTablesDev& CT = J.ct;
    const int La = J.La, Lb = J.Lb, Lc = J.Lc, Ld = J.Ld;
    const int Lab = La + Lb, Lcd = Lc + Ld, Ltot = Lab + Lcd, NS = J.NS, NGZ = J.NGZ;
    const int NTA = Lab / 2 + 1, NTC = Lcd / 2 + 1, nout = CT.nout;
    const size_t nn = (size_t)ncart * ncart;
    const int aos = J.aostride;
    double* const Bq = sm + J.oB * NB; double* const pzq = sm + J.oPz * NB; double* const Rtq = sm + J.oRt * NB;
    double* const XYq = sm + J.oXY * NB; double* const Uq = sm + J.oU * NB; double* const Sq = sm + J.oS * NB;
    double* const Itq = sm + J.oIt * NB; double* const Pstq = sm + J.oP * NB; double* const Outq = sm + J.oOut * NB;
    double* const RAq = sm + J.oRecA * NB; double* const RCq = sm + J.oRecC * NB;
    int* const aoq = reinterpret_cast<int*>(sm + J.oAO * NB);          // [NB][4][aostride]

    int qa = -1;
#pragma unroll
    for (int q = 0; q < NB; ++q) if (qa < 0 && active[q]) qa = q;
    const bool any = qa >= 0;
    const bool fill = J.fill != 0;
    const int skip = J.dbg_skip | (fill ? (32 | 64 | 128) : 0);      // fill mode: no density staging, digestion or J/K flush
    const double* recA[NB]; const double* recC[NB];
    double w[NB];
#pragma unroll
    for (int q = 0; q < NB; ++q) {
        const int ab = any ? (active[q] ? ABin[q] : ABin[qa]) : 0, cd = any ? (active[q] ? CDin[q] : CDin[qa]) : 0;
        w[q] = active[q] ? (fill ? 1.0 : win[q]) : 0.0;
        recA[q] = D.rec + D.pair_rec[ab]; recC[q] = D.rec + D.pair_rec[cd];
        if (any) {
            const int sh[4] = {D.pairA[ab], D.pairB[ab], D.pairA[cd], D.pairB[cd]};
            int* ao = aoq + q * 4 * aos;
            TUNA_LANES(x, 4 * aos) ao[x] = D.sh_ao[sh[x / aos] * SH_NCMAX + x % aos];
        }
    }
    Pol::sync();
    if (any && !(skip & 64)) {
        // stage the density blocks and clear the accumulators
        for (int dn = 0; dn < nD; ++dn) {
            const double* P = Pf + dn * nn;
            const double* Ps = Psym + dn * nn;
            TUNA_LANES(x, CT.nk) {
                const unsigned m = CT.pmap[x];
                const int ri = ((m >> 5) & 3) * aos + (m & 31), ci = ((m >> 13) & 3) * aos + ((m >> 8) & 31);
                QVec<NB> v;
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int* ao = aoq + q * 4 * aos;
                    v.v[q] = P[(size_t)ao[ri] * ncart + ao[ci]];
                }
                qst<NB>(Pstq + (size_t)(dn * nout + x) * NB, v);
            }
            TUNA_LANES(x, CT.njst) {
                QVec<NB> v;
#pragma unroll
                for (int q = 0; q < NB; ++q) v.v[q] = 0.0;
                for (unsigned t = CT.jst_ptr[x]; t < CT.jst_ptr[x + 1]; ++t) {
                    const unsigned m = CT.jst_list[t];
                    const int ri = ((m >> 5) & 3) * aos + (m & 31), ci = ((m >> 13) & 3) * aos + ((m >> 8) & 31);
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int* ao = aoq + q * 4 * aos;
                        v.v[q] += Ps[(size_t)ao[ri] * ncart + ao[ci]];
                    }
                }
                qst<NB>(Pstq + (size_t)(dn * nout + CT.nk + x) * NB, v);
            }
        }
        TUNA_LANES(x, nD * nout * NB) Outq[x] = 0.0;
    }
    const int recAsz = sp_rec_size(La, Lb), recCsz = sp_rec_size(Lc, Ld);
    const int nEzC = sp_ez_size(Lc, Ld);
    const int oExA = SP_HDR + sp_ez_size(La, Lb), oExC = SP_HDR + nEzC;
    const int nblk = (nout + 31) >> 5;

    for (int ch = 0; ch < CT.nchunk; ++ch) {
        const int e0 = CT.chunk_e0[ch], ne = CT.chunk_e0[ch + 1] - e0;
        if (any) {
            TUNA_LANES(x, ne * NB) Itq[x] = 0.0;
            if (Pol::lane() == 0) {
#pragma unroll
                for (int q = 0; q < NB; ++q) Itq[(size_t)CT.itmax * NB + q] = 0.0;
            }
        }
        for (int ia = 0; ia < J.nppAB; ++ia)
            for (int ic = 0; ic < J.nppCD; ++ic) {
                double pref[NB];
                // ---- phase 0: stage the two primitive shell-pair records (the ket z coefficients with the sign (-1)^phi folded
                // in); Boys values scaled by (-2 rho)^m and the powers of PQz, one (quartet, order) per lane -----------------
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const double* rA = recA[q] + (size_t)ia * recAsz;
                    const double* rC = recC[q] + (size_t)ic * recCsz;
                    pref[q] = 0.0;
                    if (!any || (skip & 1)) continue;
                    const double p = rA[0], qq = rC[0], pq = p + qq;
                    pref[q] = w[q] * rA[2] * rC[2] * 34.986836655249725 / (p * qq * sqrt(pq));
                    if (ic == 0) { TUNA_LANES(x, recAsz) RAq[(size_t)x * NB + q] = rA[x]; }
                    TUNA_LANES(x, recCsz) {
                        double v = rC[x];
                        const int rel = x - SP_HDR;
                        if (rel >= 0 && rel < nEzC && ((rel % (Lcd + 1)) & 1)) v = -v;
                        RCq[(size_t)x * NB + q] = v;
                    }
                }
                if (any && !(skip & 1)) {
                    TUNA_LANES(x, NB * (Ltot + 1)) {
                        const int q = x / (Ltot + 1), m = x - q * (Ltot + 1);
                        const double* rA = recA[0] + (size_t)ia * recAsz;
                        const double* rC = recC[0] + (size_t)ic * recCsz;
#pragma unroll
                        for (int k = 1; k < NB; ++k) if (q == k) { rA = recA[k] + (size_t)ia * recAsz; rC = recC[k] + (size_t)ic * recCsz; }
                        const double p = rA[0], qq = rC[0], pq = p + qq, rho = p * qq / pq, PQz = rA[1] - rC[1];
                        const double f = boys_single(D.boys, m, rho * PQz * PQz);
                        double s = 1.0, z = 1.0;
                        for (int k = 0; k < m; ++k) { s *= -2.0 * rho; z *= PQz; }
                        Bq[(size_t)m * NB + q] = f * s;
                        pzq[(size_t)m * NB + q] = z;
                    }
                }
                Pol::sync();
                // ---- phase 1: R^n_w (closed form) and the x/y convolution table ----------------------------------
                if (any && !(skip & 2)) {
                    TUNA_LANES(i, CT.n_rt) {
                        const unsigned e = CT.t_rt[i];
                        const int wv = (e >> 16) & 255, n = e >> 24;
                        QVec<NB> r;
#pragma unroll
                        for (int q = 0; q < NB; ++q) r.v[q] = 0.0;
                        for (int k = 0; 2 * k <= wv; ++k) {
                            const double h = D.herm[wv * HERM_STRIDE + k];
                            const QVec<NB> z = qld<NB>(pzq + (size_t)(wv - 2 * k) * NB), b = qld<NB>(Bq + (size_t)(n + wv - k) * NB);
#pragma unroll
                            for (int q = 0; q < NB; ++q) r.v[q] = fma(h * z.v[q], b.v[q], r.v[q]);
                        }
                        qst<NB>(Rtq + (size_t)(e & 0xffffu) * NB, r);
                    }
                    TUNA_LANES(i, CT.n_xy) {
                        const unsigned e = CT.t_xy[i];
                        const int n12 = (e >> 16) & 15, n34 = (e >> 20) & 15, m = e >> 24, px = n12 & 1;
                        const int tlo = (2 * m - n34 > px) ? 2 * m - n34 : px;
                        const int thi = (2 * m - px < n12) ? 2 * m - px : n12;
                        const double df = (n34 & 1) ? -odd_dfact(m) : odd_dfact(m);
                        const double* ExA = RAq + (size_t)(oExA + n12 * NTA) * NB;
                        const double* ExC = RCq + (size_t)(oExC + n34 * NTC) * NB;
                        QVec<NB> v;
#pragma unroll
                        for (int q = 0; q < NB; ++q) v.v[q] = 0.0;
                        for (int t = tlo; t <= thi; t += 2) {
                            const QVec<NB> a = qld<NB>(ExA + (size_t)(t >> 1) * NB), c = qld<NB>(ExC + (size_t)((2 * m - t) >> 1) * NB);
#pragma unroll
                            for (int q = 0; q < NB; ++q) v.v[q] = fma(a.v[q], c.v[q], v.v[q]);
                        }
#pragma unroll
                        for (int q = 0; q < NB; ++q) v.v[q] *= df;
                        qst<NB>(XYq + (size_t)(e & 0xffffu) * NB, v);
                    }
                }
                Pol::sync();
                // ---- phase 2: U[v][gz][n] = sum_phi (-1)^phi Ez_CD[gz][phi] R^n_{v+phi} ---------------------------
                if (any && !(skip & 4)) {
                    TUNA_LANES(i, CT.n_u) {
                        const unsigned e0w = CT.t_u[2 * i], e1w = CT.t_u[2 * i + 1];
                        const int lz34 = e1w >> 16;
                        const double* e = RCq + (size_t)(SP_HDR + (e1w & 0xffffu)) * NB;
                        const double* r = Rtq + (size_t)(e0w >> 16) * NB;
                        QVec<NB> u;
#pragma unroll
                        for (int q = 0; q < NB; ++q) u.v[q] = 0.0;
                        for (int phi = 0; phi <= lz34; ++phi) {
                            const QVec<NB> ev = qld<NB>(e + (size_t)phi * NB), rv = qld<NB>(r + (size_t)phi * NS * NB);
#pragma unroll
                            for (int q = 0; q < NB; ++q) u.v[q] = fma(ev.v[q], rv.v[q], u.v[q]);
                        }
                        qst<NB>(Uq + (size_t)(e0w & 0xffffu) * NB, u);
                    }
                }
                Pol::sync();
                // ---- phase 3: S[row][gz][n] = sum_v Ez_AB[az][bz][v] U[v][gz][n] for the chunk's bra z rows ----------
                if (any && !(skip & 8)) {
                    const int ustride = NGZ * NS;
                    for (int i = CT.chunk_s0[ch] + Pol::lane(); i < CT.chunk_s0[ch + 1]; i += Pol::G) {
                        const unsigned e0w = CT.t_s[2 * i], e1w = CT.t_s[2 * i + 1];
                        const int lz12 = e1w >> 16;
                        const double* e = RAq + (size_t)(SP_HDR + (e1w & 0xffffu)) * NB;
                        const double* u = Uq + (size_t)(e0w >> 16) * NB;
                        QVec<NB> sacc;
#pragma unroll
                        for (int q = 0; q < NB; ++q) sacc.v[q] = 0.0;
                        for (int v = 0; v <= lz12; ++v) {
                            const QVec<NB> ev = qld<NB>(e + (size_t)v * NB), uv = qld<NB>(u + (size_t)v * ustride * NB);
#pragma unroll
                            for (int q = 0; q < NB; ++q) sacc.v[q] = fma(ev.v[q], uv.v[q], sacc.v[q]);
                        }
                        qst<NB>(Sq + (size_t)(e0w & 0xffffu) * NB, sacc);
                    }
                }
                Pol::sync();
                // ---- phase 4: table-driven integral assembly, accumulated over primitive quartets; the table entry of the
                // lane's next integral is fetched while the current one is assembled -------------------------------------
                if (any && !(skip & 16)) {
                    const uint2* p4 = reinterpret_cast<const uint2*>(CT.p4) + e0;
                    int e = Pol::lane();
                    uint2 nxt = make_uint2(0u, 0u);
                    if (e < ne) nxt = p4[e];
                    for (; e < ne; e += Pol::G) {
                        const unsigned w0 = nxt.x, w1 = nxt.y;
                        if (e + Pol::G < ne) nxt = p4[e + Pol::G];
                        const int xo = w0 & 0xffffu, yo = w0 >> 16, so = w1 & 0xffffu;
                        const int mx0 = (w1 >> 16) & 15, mx1 = (w1 >> 20) & 15, my0 = (w1 >> 24) & 15, my1 = (w1 >> 28) & 15;
                        QVec<NB> val;
#pragma unroll
                        for (int q = 0; q < NB; ++q) val.v[q] = 0.0;
                        const double* xyx = XYq + (size_t)xo * NB;
                        const double* xyy = XYq + (size_t)yo * NB;
#ifdef TUNA_SHELL_ASM_UNROLL
                        const double* s_row = Sq + (size_t)so * NB;
                        switch (my1 - my0) {
                            case 0: val = assemble_integral<NB, 1>(xyx, xyy, s_row, mx0, mx1, my0); break;
                            case 1: val = assemble_integral<NB, 2>(xyx, xyy, s_row, mx0, mx1, my0); break;
                            case 2: val = assemble_integral<NB, 3>(xyx, xyy, s_row, mx0, mx1, my0); break;
                            case 3: val = assemble_integral<NB, 4>(xyx, xyy, s_row, mx0, mx1, my0); break;
                            default:
#endif
                        for (int m = mx0; m <= mx1; ++m) {
                            QVec<NB> t;
#pragma unroll
                            for (int q = 0; q < NB; ++q) t.v[q] = 0.0;
                            const double* sp = Sq + (size_t)(so + m) * NB;
                            for (int mp = my0; mp <= my1; ++mp) {
                                const QVec<NB> y = qld<NB>(xyy + (size_t)mp * NB), s = qld<NB>(sp + (size_t)mp * NB);
#pragma unroll
                                for (int q = 0; q < NB; ++q) t.v[q] = fma(y.v[q], s.v[q], t.v[q]);
                            }
                            const QVec<NB> x = qld<NB>(xyx + (size_t)m * NB);
#pragma unroll
                            for (int q = 0; q < NB; ++q) val.v[q] = fma(x.v[q], t.v[q], val.v[q]);
                        }
#ifdef TUNA_SHELL_ASM_UNROLL
                        }
#endif
                        QVec<NB> it = qld<NB>(Itq + (size_t)e * NB);
#pragma unroll
                        for (int q = 0; q < NB; ++q) it.v[q] = fma(pref[q], val.v[q], it.v[q]);
                        qst<NB>(Itq + (size_t)e * NB, it);
                    }
                }
            }
        Pol::sync();
        // ---- phase 5: table-driven digestion of the chunk: every accumulator is owned by one lane.  The term lists of 32
        // consecutive accumulators are stored transposed (term quad t of accumulator o at (t * 32 + (o & 31))), padded to the
        // longest list of the block with dummy terms that read the zero slot It[itmax]: one coalesced 16-byte load per lane
        // brings four terms, and the next quad is in flight while the current one is digested.
        if (any && fill) {
            // ---- fill mode: scatter the chunk's integrals (normalised) to the eight images of every CANONICAL AO quartet
            // (i >= j, k >= l, ij >= kl when shells coincide), so the dense tensor is exactly 8-fold symmetric and deterministic
            const int f0 = CT.chunk_f0[ch], nf = CT.chunk_f0[ch + 1] - f0;
            const uint2* fl = reinterpret_cast<const uint2*>(CT.p6) + f0;
            const int ncB = (Lb + 1) * (Lb + 2) / 2, ncD = (Ld + 1) * (Ld + 2) / 2;
            const size_t n1 = (size_t)ncart, n2 = n1 * n1, n3 = n2 * n1;
            bool sameAB[NB], sameCD[NB], diag[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int ab = active[q] ? ABin[q] : 0, cd = active[q] ? CDin[q] : 0;
                sameAB[q] = D.pairA[ab] == D.pairB[ab]; sameCD[q] = D.pairA[cd] == D.pairB[cd]; diag[q] = ab == cd;
            }
            TUNA_LANES(x, nf) {
                const uint2 ent = fl[x];
                const int slot = ent.x & 0xffffu, a = (ent.x >> 16) & 31, b = (ent.x >> 21) & 31, c = (ent.x >> 26) & 31, d = (int)ent.y;
                const QVec<NB> v = qld<NB>(Itq + (size_t)slot * NB);
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    if (!active[q]) continue;
                    if ((sameAB[q] && b > a) || (sameCD[q] && d > c) || (diag[q] && c * ncD + d > a * ncB + b)) continue;
                    const int* ao = aoq + q * 4 * aos;
                    const size_t i = ao[a], j = ao[aos + b], k = ao[2 * aos + c], l = ao[3 * aos + d];
                    const double val = v.v[q] * D.fnorm[i] * D.fnorm[j] * D.fnorm[k] * D.fnorm[l];
                    double* E = D.eri_out;
                    E[i * n3 + j * n2 + k * n1 + l] = val; E[j * n3 + i * n2 + k * n1 + l] = val;
                    E[i * n3 + j * n2 + l * n1 + k] = val; E[j * n3 + i * n2 + l * n1 + k] = val;
                    E[k * n3 + l * n2 + i * n1 + j] = val; E[l * n3 + k * n2 + i * n1 + j] = val;
                    E[k * n3 + l * n2 + j * n1 + i] = val; E[l * n3 + k * n2 + j * n1 + i] = val;
                }
            }
        }
#ifdef TUNA_SHELL_WIDE_TERMS
        if (any && !(skip & 32)) {
            const unsigned* ptr = CT.p5ptr + (size_t)ch * (nblk + 1);
            const uint4* term = reinterpret_cast<const uint4*>(CT.p5w + 2 * (size_t)CT.p5off[ch]);
            TUNA_LANES(o, nout) {
                const unsigned b0 = ptr[o >> 5], nq = (ptr[(o >> 5) + 1] - b0) >> 5;
                if (nq == 0) continue;
                for (int dn = 0; dn < nD; ++dn) {
                    // both byte offsets of a term are relative to the It buffer (the density offsets carry oP - oIt), so every
                    // operand address is one base register plus the table word
                    const char* base = reinterpret_cast<const char*>(Itq) + (size_t)dn * nout * NB * sizeof(double);
                    const uint4* tp = term + 2 * (size_t)b0 + (o & 31);      // half h of quad t at tp[(2 t + h) * 32]
                    double s0[NB], s1[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) { s0[q] = 0.0; s1[q] = 0.0; }
#define TUNA_DIGEST_QUAD(TA, TB)                                                                                                   \
    {                                                                                                                              \
        const QVec<NB> i0 = qld<NB>(reinterpret_cast<const double*>(base - (size_t)dn * nout * NB * sizeof(double) + (TA).x));     \
        const QVec<NB> p0 = qld<NB>(reinterpret_cast<const double*>(base + (TA).y));                                               \
        const QVec<NB> i1 = qld<NB>(reinterpret_cast<const double*>(base - (size_t)dn * nout * NB * sizeof(double) + (TA).z));     \
        const QVec<NB> p1 = qld<NB>(reinterpret_cast<const double*>(base + (TA).w));                                               \
        const QVec<NB> i2 = qld<NB>(reinterpret_cast<const double*>(base - (size_t)dn * nout * NB * sizeof(double) + (TB).x));     \
        const QVec<NB> p2 = qld<NB>(reinterpret_cast<const double*>(base + (TB).y));                                               \
        const QVec<NB> i3 = qld<NB>(reinterpret_cast<const double*>(base - (size_t)dn * nout * NB * sizeof(double) + (TB).z));     \
        const QVec<NB> p3 = qld<NB>(reinterpret_cast<const double*>(base + (TB).w));                                               \
        _Pragma("unroll") for (int q = 0; q < NB; ++q) {                                                                           \
            s0[q] = fma(i0.v[q], p0.v[q], s0[q]);                                                                                  \
            s1[q] = fma(i1.v[q], p1.v[q], s1[q]);                                                                                  \
            s0[q] = fma(i2.v[q], p2.v[q], s0[q]);                                                                                  \
            s1[q] = fma(i3.v[q], p3.v[q], s1[q]);                                                                                  \
        }                                                                                                                          \
    }
                    // two quads per trip in two register sets (no copies); the loads of the next trip are issued before the FMAs
                    uint4 a0 = tp[0], a1 = tp[32];
                    unsigned t = 0;
                    for (; t + 1 < nq; t += 2) {
                        const uint4 c0 = tp[64], c1 = tp[96];
                        TUNA_DIGEST_QUAD(a0, a1)
                        tp += 128;
                        if (t + 2 < nq) { a0 = tp[0]; a1 = tp[32]; }
                        TUNA_DIGEST_QUAD(c0, c1)
                    }
                    if (t < nq) TUNA_DIGEST_QUAD(a0, a1)
#undef TUNA_DIGEST_QUAD
                    QVec<NB> out = qld<NB>(Outq + (size_t)(dn * nout + o) * NB);
#pragma unroll
                    for (int q = 0; q < NB; ++q) out.v[q] += s0[q] + s1[q];
                    qst<NB>(Outq + (size_t)(dn * nout + o) * NB, out);
                }
            }
        }
#else
        if (any && !(skip & 32)) {
            const unsigned* ptr = CT.p5ptr + (size_t)ch * (nblk + 1);
            const uint4* term = reinterpret_cast<const uint4*>(CT.p5term + CT.p5off[ch]);
            TUNA_LANES(o, nout) {
                const unsigned b0 = ptr[o >> 5], nq = (ptr[(o >> 5) + 1] - b0) >> 5;
                if (nq == 0) continue;
                const uint4* tp = term + b0 + (o & 31);
                for (int dn = 0; dn < nD; ++dn) {
                    const double* Pd = Pstq + (size_t)dn * nout * NB;
                    double s0[NB], s1[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) { s0[q] = 0.0; s1[q] = 0.0; }
                    uint4 nxt = tp[0];
                    for (unsigned t = 0; t < nq; ++t) {
                        const uint4 tq = nxt;
                        if (t + 1 < nq) nxt = tp[(size_t)(t + 1) * 32];
                        const QVec<NB> i0 = qld<NB>(Itq + (size_t)(tq.x & 0xffffu) * NB), p0 = qld<NB>(Pd + (size_t)(tq.x >> 16) * NB);
                        const QVec<NB> i1 = qld<NB>(Itq + (size_t)(tq.y & 0xffffu) * NB), p1 = qld<NB>(Pd + (size_t)(tq.y >> 16) * NB);
                        const QVec<NB> i2 = qld<NB>(Itq + (size_t)(tq.z & 0xffffu) * NB), p2 = qld<NB>(Pd + (size_t)(tq.z >> 16) * NB);
                        const QVec<NB> i3 = qld<NB>(Itq + (size_t)(tq.w & 0xffffu) * NB), p3 = qld<NB>(Pd + (size_t)(tq.w >> 16) * NB);
#pragma unroll
                        for (int q = 0; q < NB; ++q) {
                            s0[q] = fma(i0.v[q], p0.v[q], s0[q]);
                            s1[q] = fma(i1.v[q], p1.v[q], s1[q]);
                            s0[q] = fma(i2.v[q], p2.v[q], s0[q]);
                            s1[q] = fma(i3.v[q], p3.v[q], s1[q]);
                        }
                    }
                    QVec<NB> out = qld<NB>(Outq + (size_t)(dn * nout + o) * NB);
#pragma unroll
                    for (int q = 0; q < NB; ++q) out.v[q] += s0[q] + s1[q];
                    qst<NB>(Outq + (size_t)(dn * nout + o) * NB, out);
                }
            }
        }
#endif
        Pol::sync();
    }
    // ---- flush the shell blocks: one atomic per block entry per shell quartet ---------------------------------------
    if (any && !(skip & 128)) {
        for (int dn = 0; dn < nD; ++dn) {
            TUNA_LANES(x, nout) {
                const unsigned m = CT.omap[x];
                if (m == 0xffffu) continue;                    // J pair-function accumulator: expanded below
                const int ri = ((m >> 5) & 3) * aos + (m & 31), ci = ((m >> 13) & 3) * aos + ((m >> 8) & 31);
                const QVec<NB> v = qld<NB>(Outq + (size_t)(dn * nout + x) * NB);
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    if (!active[q]) continue;
                    const int* ao = aoq + q * 4 * aos;
                    Pol::atomic_add(Kf + dn * nn + (size_t)ao[ri] * ncart + ao[ci], v.v[q]);
                }
            }
            TUNA_LANES(x, CT.njfl) {
                const unsigned e = CT.jflush[x], m = e & 0xffffu;
                const int ri = ((m >> 5) & 3) * aos + (m & 31), ci = ((m >> 13) & 3) * aos + ((m >> 8) & 31);
                const QVec<NB> v = qld<NB>(Outq + (size_t)(dn * nout + (e >> 16)) * NB);
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    if (!active[q]) continue;
                    const int* ao = aoq + q * 4 * aos;
                    Pol::atomic_add(Jf + dn * nn + (size_t)ao[ri] * ncart + ao[ci], v.v[q]);
                }
            }
        }
    }
    Pol::sync();
}

// Single-quartet convenience wrapper.
template <class Pol>
TUNA_HD void shell_quartet(const ShellJob& J, const ShellData& D, bool active, int AB, int CD, double w, double* __restrict__ sm,
                           int nD, const double* __restrict__ Pf, const double* __restrict__ Psym, double* Jf, double* Kf, int ncart) {
    shell_quartets<Pol, 1>(J, D, &active, &AB, &CD, &w, sm, nD, Pf, Psym, Jf, Kf, ncart);
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
