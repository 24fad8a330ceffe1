// shell_host.hpp — host-side preparation for the shell-quartet engine (shell_jk.cuh): grouping the reference's
// per-component Basis list into shells, the primitive shell-pair records, pair classes and the job list.
// Built once per geometry.  Replaces, at shell granularity, the pair cache of
// TUNA/tuna_integrals/tuna_integral.pyx:1050-1128 and the pair12 >= pair34 double loop of :1312-1331.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <tuple>
#include <vector>

#include "pairtable.hpp"
#include "shell_jk.cuh"

namespace tuna {

struct HostShell {
    int L = 0, nprim = 0;
    double z = 0.0;
    std::vector<double> exps, coef;     // shell-level coefficients: ceff / f(component)
    int ao[SH_NCMAX];
    int filled = 0;
};

struct ShellSystem {
    bool ok = false;
    std::vector<HostShell> shells;
    std::vector<double> fnorm;           // per AO: f = 1 / sqrt((2l-1)!! (2m-1)!! (2n-1)!!)
    std::vector<int> sh_ao;              // [shell * SH_NCMAX + c]
    // shell pairs
    std::vector<int> pairA, pairB;
    std::vector<long long> pair_rec;
    std::vector<double> rec;
    std::vector<double> pairQ;
    // classes: key (La, Lb, npp)
    struct PairClassInfo { int La, Lb, npp; std::vector<int> pairs; };
    std::vector<PairClassInfo> classes;
};

inline double dfact_odd(int n) { double r = 1.0; while (n > 1) { r *= n; n -= 2; } return r; }

// Group the per-component list into full Cartesian shells.  Components of one shell need not be contiguous
// (DECONTRACT emits component-major order, tuna_molecule.py:557-565).  Returns false if the list is not a union of
// complete shells with a common radial part; the caller then uses the per-component kernels.
inline bool detect_shells(const HostBasis& B, const ShellTab& T, ShellSystem& S) {
    S = ShellSystem();
    const int n = B.ncart;
    S.fnorm.resize(n);
    std::map<std::tuple<int, int, int>, int> comp_index[SH_LMAX + 1];
    for (int L = 0; L <= SH_LMAX; ++L)
        for (int c = 0; c < T.nc[L]; ++c) comp_index[L][{T.lx[L][c], T.ly[L][c], T.lz[L][c]}] = c;
    std::map<std::tuple<double, int, int>, std::vector<int>> buckets;     // (z, L, nprim) -> candidate shells
    for (int i = 0; i < n; ++i) {
        const int l = B.lmn[3 * i], m = B.lmn[3 * i + 1], nn = B.lmn[3 * i + 2], L = l + m + nn;
        const double f = 1.0 / std::sqrt(dfact_odd(2 * l - 1) * dfact_odd(2 * m - 1) * dfact_odd(2 * nn - 1));
        S.fnorm[i] = f;
        const int c = comp_index[L][{l, m, nn}];
        const int np = B.nprim[i];
        auto& cand = buckets[{B.oz[i], L, np}];
        int found = -1;
        for (int s : cand) {
            HostShell& sh = S.shells[s];
            if (sh.ao[c] >= 0) continue;
            bool same = true;
            for (int k = 0; k < np && same; ++k) {
                const double ck = B.ceff[B.off[i] + k] / f;
                same = sh.exps[k] == B.exps[B.off[i] + k] && std::fabs(ck - sh.coef[k]) <= 1e-11 * std::fabs(sh.coef[k]);
            }
            if (same) { found = s; break; }
        }
        if (found < 0) {
            HostShell sh;
            sh.L = L; sh.nprim = np; sh.z = B.oz[i];
            for (int k = 0; k < np; ++k) { sh.exps.push_back(B.exps[B.off[i] + k]); sh.coef.push_back(B.ceff[B.off[i] + k] / f); }
            for (int k = 0; k < SH_NCMAX; ++k) sh.ao[k] = -1;
            found = (int)S.shells.size();
            S.shells.push_back(sh);
            cand.push_back(found);
        }
        S.shells[found].ao[c] = i;
        S.shells[found].filled++;
    }
    for (const HostShell& sh : S.shells)
        if (sh.filled != T.nc[sh.L]) return false;
    S.sh_ao.assign(S.shells.size() * SH_NCMAX, 0);
    for (size_t s = 0; s < S.shells.size(); ++s)
        for (int c = 0; c < T.nc[S.shells[s].L]; ++c) S.sh_ao[s * SH_NCMAX + c] = S.shells[s].ao[c];
    S.ok = true;
    return true;
}

// AO-level Schwarz factors (from the per-component pair table, already on the host) -> shell pairs, records, classes.
inline void build_shell_pairs(ShellSystem& S, const ShellTab& T, const PairTable& PT, const std::vector<double>& aoQ, int ncart) {
    const int ns = (int)S.shells.size();
    std::vector<int> lookup((size_t)ncart * ncart, -1);
    for (int64_t a = 0; a < PT.npair; ++a) {
        lookup[(size_t)PT.pi[a] * ncart + PT.pj[a]] = (int)a;
        lookup[(size_t)PT.pj[a] * ncart + PT.pi[a]] = (int)a;
    }
    S.pairA.clear(); S.pairB.clear(); S.pair_rec.clear(); S.pairQ.clear(); S.classes.clear();
    std::map<std::tuple<int, int, int>, int> class_of;
    long long rec_total = 0;
    for (int a = 0; a < ns; ++a)
        for (int b = 0; b <= a; ++b) {
            int A = a, Bs = b;
            if (S.shells[A].L < S.shells[Bs].L) std::swap(A, Bs);
            const HostShell& sa = S.shells[A];
            const HostShell& sb = S.shells[Bs];
            const int npp = sa.nprim * sb.nprim;
            double q = 0.0;
            for (int ca = 0; ca < T.nc[sa.L]; ++ca)
                for (int cb = 0; cb < T.nc[sb.L]; ++cb) q = std::max(q, aoQ[lookup[(size_t)sa.ao[ca] * ncart + sb.ao[cb]]]);
            const int id = (int)S.pairA.size();
            S.pairA.push_back(A); S.pairB.push_back(Bs); S.pairQ.push_back(q);
            S.pair_rec.push_back(rec_total);
            rec_total += (long long)npp * sp_rec_size(sa.L, sb.L);
            auto key = std::make_tuple(sa.L, sb.L, npp);
            auto it = class_of.find(key);
            if (it == class_of.end()) {
                it = class_of.emplace(key, (int)S.classes.size()).first;
                S.classes.push_back({sa.L, sb.L, npp, {}});
            }
            S.classes[it->second].pairs.push_back(id);
        }
    // deterministic class order: light to heavy; Schwarz-descending inside a class
    std::sort(S.classes.begin(), S.classes.end(), [](const ShellSystem::PairClassInfo& x, const ShellSystem::PairClassInfo& y) {
        return std::make_tuple(x.La + x.Lb, x.La, x.npp) < std::make_tuple(y.La + y.Lb, y.La, y.npp);
    });
    for (auto& c : S.classes)
        std::stable_sort(c.pairs.begin(), c.pairs.end(), [&](int x, int y) { return S.pairQ[x] > S.pairQ[y]; });
    S.rec.assign((size_t)rec_total, 0.0);
    const int np = (int)S.pairA.size();
#pragma omp parallel
    {
        std::vector<double> scratch;
#pragma omp for schedule(dynamic, 64)
        for (int id = 0; id < np; ++id) {
            const HostShell& sa = S.shells[S.pairA[id]];
            const HostShell& sb = S.shells[S.pairB[id]];
            const int La = sa.L, Lb = sb.L, Lab = La + Lb, rsz = sp_rec_size(La, Lb), nez = sp_ez_size(La, Lb), NT = Lab / 2 + 1;
            double* r = S.rec.data() + S.pair_rec[id];
            for (int ka = 0; ka < sa.nprim; ++ka)
                for (int kb = 0; kb < sb.nprim; ++kb, r += rsz) {
                    const double a = sa.exps[ka], b = sb.exps[kb], p = a + b;
                    r[0] = p; r[1] = (a * sa.z + b * sb.z) / p; r[2] = sa.coef[ka] * sb.coef[kb]; r[3] = 0.0;
                    hermite_table(La, Lb, sa.z - sb.z, a, b, scratch);
                    for (int i = 0; i <= La; ++i)
                        for (int j = 0; j <= Lb; ++j)
                            for (int t = 0; t <= i + j; ++t)
                                r[SP_HDR + (i * (Lb + 1) + j) * (Lab + 1) + t] = scratch[((size_t)i * (Lb + 1) + j) * (Lab + 2) + t];
                    hermite_table(Lab, 0, 0.0, a, b, scratch);
                    for (int nn = 0; nn <= Lab; ++nn)
                        for (int t = nn & 1; t <= nn; t += 2) r[SP_HDR + nez + nn * NT + (t >> 1)] = scratch[(size_t)nn * (Lab + 2) + t];
                }
        }
    }
}

// Kets kept per bra for one (bra class, ket class) job: Schwarz cut on the descending ket list; same class -> ket_pos <= bra_pos.
inline long long build_item_prefix(const ShellSystem& S, int cb, int ck, double tau_static, std::vector<long long>& prefix) {
    const auto& bra = S.classes[cb].pairs;
    const auto& ket = S.classes[ck].pairs;
    prefix.assign(bra.size() + 1, 0);
    for (size_t i = 0; i < bra.size(); ++i) {
        const double qb = S.pairQ[bra[i]];
        size_t lo = 0, hi = ket.size();       // first ket with Q_k * qb < tau_static
        if (tau_static > 0.0)
            while (lo < hi) {
                const size_t mid = (lo + hi) / 2;
                if (S.pairQ[ket[mid]] * qb >= tau_static) lo = mid + 1; else hi = mid;
            }
        else
            lo = ket.size();
        size_t keep = lo;
        if (cb == ck) keep = std::min(keep, i + 1);
        prefix[i + 1] = prefix[i] + (long long)keep;
    }
    return prefix.back();
}

// ---------------------------------------------------------------------------------------------------------------
// Per-class work tables of the shell-quartet engine (see ClassTablesDev in shell_jk.cuh).
// ---------------------------------------------------------------------------------------------------------------
struct ClassTablesHost {
    int La, Lb, Lc, Ld;
    int nout = 0, nk = 0, itmax = 0, smax_rows = 0, nint = 0;
    std::vector<int> chunk_bz0, chunk_e0, chunk_s0, chunk_f0, bz_list;
    std::vector<unsigned> p4, p5ptr, p5term, p5off, p6, t_rt, t_xy, t_u, t_s;
    std::vector<unsigned short> pmap, omap, jst_list;
    std::vector<unsigned> jst_ptr, jflush;
#ifdef TUNA_SHELL_WIDE_TERMS
    std::vector<unsigned> p5wide;     // phase-5 terms as (It slot, staged density entry) word pairs in ELEMENT units; scale_wide_terms() makes the device copy
#endif
    long long p5real = 0;             // digestion terms before padding (table statistics)
    long long allowed = 0;            // parity-allowed component quartets = integrals per shell quartet
    double uniq[6] = {0, 0, 0, 0, 0, 0};   // unique AO quartets a shell quartet stands for: [0] generic, [1] A==B, [2] C==D,
                                           // [3] A==B and C==D, [4] AB==CD (A!=B), [5] all four shells equal
};

constexpr int SH_IT_BUDGET = 6144;    // doubles of shared memory for the integral buffer of a chunk
constexpr int SH_S_BUDGET = 6144;     // doubles for the S slice of a chunk

inline void build_class_tables(const ShellTab& T, int La, int Lb, int Lc, int Ld, ClassTablesHost& C, int it_budget = SH_IT_BUDGET,
                               int s_budget = SH_S_BUDGET, bool with_fill = false) {
    C = ClassTablesHost();
    C.La = La; C.Lb = Lb; C.Lc = Lc; C.Ld = Ld;
    const int Lab = La + Lb, Lcd = Lc + Ld, Ltot = Lab + Lcd, NS = Ltot / 2 + 1, NGZ = (Lc + 1) * (Ld + 1);
    const int ncA = T.nc[La], ncB = T.nc[Lb], ncC = T.nc[Lc], ncD = T.nc[Ld];
    // K: output blocks KAC KAD KBC KBD and staged density blocks PDB PCB PDA PCA, one entry per component pair.
    // J: accumulated per bra / ket PAIR FUNCTION (ax+bx, ay+by, az, bz): accumulators Jb[beta], Jg[gamma] and staged
    //    symmetrised densities Pg[gamma] = sum_{(c,d)->gamma} Psym[c][d], Pb[beta] = sum_{(a,b)->beta} Psym[a][b].
    const int ob[5] = {0, ncA * ncC, ncA * ncC + ncA * ncD, ncA * ncC + ncA * ncD + ncB * ncC, ncA * ncC + ncA * ncD + ncB * ncC + ncB * ncD};
    C.nk = ob[4];
    const int pb[4] = {0, ncD * ncB, ncD * ncB + ncC * ncB, ncD * ncB + ncC * ncB + ncD * ncA};
    auto rc = [](int rsel, int r, int csel, int c) { return (unsigned short)((rsel << 5) | r | ((csel << 5 | c) << 8)); };
    std::map<std::tuple<int, int, int, int>, int> beta_of, gamma_of;          // (nx, ny, z1, z2) -> pair-function index
    std::vector<int> bidx(ncA * ncB), gidx(ncC * ncD);
    for (int a = 0; a < ncA; ++a)
        for (int b = 0; b < ncB; ++b) {
            auto key = std::make_tuple(T.lx[La][a] + T.lx[Lb][b], T.ly[La][a] + T.ly[Lb][b], T.lz[La][a], T.lz[Lb][b]);
            auto it = beta_of.find(key);
            if (it == beta_of.end()) it = beta_of.emplace(key, (int)beta_of.size()).first;
            bidx[a * ncB + b] = it->second;
        }
    for (int c = 0; c < ncC; ++c)
        for (int d = 0; d < ncD; ++d) {
            auto key = std::make_tuple(T.lx[Lc][c] + T.lx[Ld][d], T.ly[Lc][c] + T.ly[Ld][d], T.lz[Lc][c], T.lz[Ld][d]);
            auto it = gamma_of.find(key);
            if (it == gamma_of.end()) it = gamma_of.emplace(key, (int)gamma_of.size()).first;
            gidx[c * ncD + d] = it->second;
        }
    const int nbeta = (int)beta_of.size(), ngamma = (int)gamma_of.size();
    // accumulator ids (before sorting): [0, nk) K entries, [nk, nk+nbeta) Jb, [nk+nbeta, nk+nbeta+ngamma) Jg
    // staged density ids:               [0, nk) K entries, [nk, nk+ngamma) Pg, [nk+ngamma, nk+ngamma+nbeta) Pb
    C.nout = C.nk + nbeta + ngamma;
    C.omap.assign(C.nout, 0xffff); C.pmap.assign(C.nk, 0);
    for (int a = 0; a < ncA; ++a) for (int c = 0; c < ncC; ++c) C.omap[ob[0] + a * ncC + c] = rc(0, a, 2, c);
    for (int a = 0; a < ncA; ++a) for (int d = 0; d < ncD; ++d) C.omap[ob[1] + a * ncD + d] = rc(0, a, 3, d);
    for (int b = 0; b < ncB; ++b) for (int c = 0; c < ncC; ++c) C.omap[ob[2] + b * ncC + c] = rc(1, b, 2, c);
    for (int b = 0; b < ncB; ++b) for (int d = 0; d < ncD; ++d) C.omap[ob[3] + b * ncD + d] = rc(1, b, 3, d);
    for (int d = 0; d < ncD; ++d) for (int b = 0; b < ncB; ++b) C.pmap[pb[0] + d * ncB + b] = rc(3, d, 1, b);
    for (int c = 0; c < ncC; ++c) for (int b = 0; b < ncB; ++b) C.pmap[pb[1] + c * ncB + b] = rc(2, c, 1, b);
    for (int d = 0; d < ncD; ++d) for (int a = 0; a < ncA; ++a) C.pmap[pb[2] + d * ncA + a] = rc(3, d, 0, a);
    for (int c = 0; c < ncC; ++c) for (int a = 0; a < ncA; ++a) C.pmap[pb[3] + c * ncA + a] = rc(2, c, 0, a);
    {   // staging CSR of the pair-function densities: first the gammas, then the betas
        std::vector<std::vector<unsigned short>> lists(ngamma + nbeta);
        for (int c = 0; c < ncC; ++c) for (int d = 0; d < ncD; ++d) lists[gidx[c * ncD + d]].push_back(rc(2, c, 3, d));
        for (int a = 0; a < ncA; ++a) for (int b = 0; b < ncB; ++b) lists[ngamma + bidx[a * ncB + b]].push_back(rc(0, a, 1, b));
        C.jst_ptr.push_back(0);
        for (auto& l : lists) { C.jst_list.insert(C.jst_list.end(), l.begin(), l.end()); C.jst_ptr.push_back((unsigned)C.jst_list.size()); }
    }

    // phase 1-2 work lists
    for (int w = 0; w <= Ltot; ++w)
        for (int n = 0; 2 * n + w <= Ltot; ++n) C.t_rt.push_back((unsigned)(w * NS + n) | (unsigned)w << 16 | (unsigned)n << 24);
    for (int n12 = 0; n12 <= Lab; ++n12)
        for (int n34 = n12 & 1; n34 <= Lcd; n34 += 2)
            for (int m = n12 & 1; 2 * m <= n12 + n34; ++m)
                C.t_xy.push_back((unsigned)((n12 * (Lcd + 1) + n34) * NS + m) | (unsigned)n12 << 16 | (unsigned)n34 << 20 | (unsigned)m << 24);
    for (int v = 0; v <= Lab; ++v)
        for (int cz = 0; cz <= Lc; ++cz)
            for (int dz = 0; dz <= Ld; ++dz) {
                const int gz = cz * (Ld + 1) + dz, lz34 = cz + dz;
                for (int n = 0; 2 * n + v + lz34 <= Ltot; ++n) {
                    C.t_u.push_back((unsigned)((v * NGZ + gz) * NS + n) | (unsigned)(v * NS + n) << 16);
                    C.t_u.push_back((unsigned)(gz * (Lcd + 1)) | (unsigned)lz34 << 16);
                }
            }

    // divergence control: lanes of a warp take consecutive entries, so order every work list by inner-loop length
    auto sort_pairs = [](std::vector<unsigned>& v, size_t lo, size_t hi, auto keyfn) {      // v holds 2-word entries in [lo, hi)
        std::vector<std::pair<unsigned, unsigned>> tmp;
        for (size_t i = lo; i < hi; i += 2) tmp.push_back({v[i], v[i + 1]});
        std::stable_sort(tmp.begin(), tmp.end(), [&](const auto& x, const auto& y) { return keyfn(x) > keyfn(y); });
        for (size_t i = 0; i < tmp.size(); ++i) { v[lo + 2 * i] = tmp[i].first; v[lo + 2 * i + 1] = tmp[i].second; }
    };
    std::stable_sort(C.t_rt.begin(), C.t_rt.end(), [](unsigned x, unsigned y) { return ((x >> 16) & 255) > ((y >> 16) & 255); });
    sort_pairs(C.t_u, 0, C.t_u.size(), [](const std::pair<unsigned, unsigned>& e) { return e.second >> 16; });

    // bra z-combinations, chunked so that the S slice and the integral buffer fit their budgets
    struct Quartet { int a, b, c, d; };
    std::vector<std::vector<Quartet>> per_bz;
    for (int az = 0; az <= La; ++az)
        for (int bz = 0; bz <= Lb; ++bz) {
            C.bz_list.push_back(az | bz << 8);
            std::vector<Quartet> q;
            for (int a = 0; a < ncA; ++a) {
                if (T.lz[La][a] != az) continue;
                for (int b = 0; b < ncB; ++b) {
                    if (T.lz[Lb][b] != bz) continue;
                    for (int c = 0; c < ncC; ++c)
                        for (int d = 0; d < ncD; ++d)
                            if ((T.pg[La][a] ^ T.pg[Lb][b] ^ T.pg[Lc][c] ^ T.pg[Ld][d]) == 0) q.push_back({a, b, c, d});
                }
            }
            per_bz.push_back(q);
        }
    const int nbz = (int)C.bz_list.size();
    const int max_rows = std::max(1, std::min(s_budget / (NGZ * NS), 65535 / (NGZ * NS)));
    C.chunk_bz0.push_back(0); C.chunk_e0.push_back(0); C.chunk_s0.push_back(0);
    int cur_rows = 0, cur_int = 0;
    auto pf_key = [&](int bi, const Quartet& q) {
        const int nx12 = T.lx[La][q.a] + T.lx[Lb][q.b], nx34 = T.lx[Lc][q.c] + T.lx[Ld][q.d];
        const int ny12 = T.ly[La][q.a] + T.ly[Lb][q.b], ny34 = T.ly[Lc][q.c] + T.ly[Ld][q.d];
        const int gz = T.lz[Lc][q.c] * (Ld + 1) + T.lz[Ld][q.d];
        return ((((long long)bi * 16 + nx12) * 16 + ny12) * 16 + nx34) * 16 * 64 + ny34 * 64 + gz;
    };
    std::map<long long, int> slot_of;     // pair-function quartet -> It slot inside its chunk
    for (int bi = 0; bi < nbz; ++bi) {
        int distinct = 0;
        {
            std::map<long long, int> seen;
            for (const Quartet& q : per_bz[bi]) seen[pf_key(bi, q)] = 1;
            distinct = (int)seen.size();
        }
        const size_t slots_before = slot_of.size();
        if (cur_rows > 0 && (cur_rows + 1 > max_rows || cur_int + distinct > it_budget)) {
            C.chunk_bz0.push_back(bi); C.chunk_e0.push_back(C.nint); C.chunk_s0.push_back((int)C.t_s.size() / 2);
            cur_rows = 0; cur_int = 0;
        }
        const int az = C.bz_list[bi] & 255, bz = C.bz_list[bi] >> 8, lz12 = az + bz;
        // phase 3 entries of this row
        for (int cz = 0; cz <= Lc; ++cz)
            for (int dz = 0; dz <= Ld; ++dz) {
                const int gz = cz * (Ld + 1) + dz;
                for (int n = 0; 2 * n + lz12 + cz + dz <= Ltot; ++n) {
                    C.t_s.push_back((unsigned)((cur_rows * NGZ + gz) * NS + n) | (unsigned)(gz * NS + n) << 16);
                    C.t_s.push_back((unsigned)((az * (Lb + 1) + bz) * (Lab + 1)) | (unsigned)lz12 << 16);
                }
            }
        // phase 4 entries: one per DISTINCT pair-function quartet.  Component quartets with equal (ax+bx, ay+by, az, bz |
        // cx+dx, cy+dy, cz, dz) have identical (unnormalised) integrals, so they share one It slot (1.8x-3.2x fewer).
        for (const Quartet& q : per_bz[bi]) {
            const int nx12 = T.lx[La][q.a] + T.lx[Lb][q.b], nx34 = T.lx[Lc][q.c] + T.lx[Ld][q.d];
            const int ny12 = T.ly[La][q.a] + T.ly[Lb][q.b], ny34 = T.ly[Lc][q.c] + T.ly[Ld][q.d];
            const int gz = T.lz[Lc][q.c] * (Ld + 1) + T.lz[Ld][q.d];
            const long long key = ((((long long)bi * 16 + nx12) * 16 + ny12) * 16 + nx34) * 16 * 64 + ny34 * 64 + gz;
            if (slot_of.count(key)) continue;
            slot_of[key] = cur_int + (int)(slot_of.size() - slots_before);
            const unsigned xoff = (nx12 * (Lcd + 1) + nx34) * NS, yoff = (ny12 * (Lcd + 1) + ny34) * NS, soff = (cur_rows * NGZ + gz) * NS;
            C.p4.push_back(xoff | yoff << 16);
            C.p4.push_back(soff | (unsigned)(nx12 & 1) << 16 | (unsigned)((nx12 + nx34) >> 1) << 20 | (unsigned)(ny12 & 1) << 24 | (unsigned)((ny12 + ny34) >> 1) << 28);
        }
        const int add = (int)(slot_of.size() - slots_before);
        C.nint += add; cur_int += add; cur_rows += 1;
        C.itmax = std::max(C.itmax, cur_int);
        C.smax_rows = std::max(C.smax_rows, cur_rows);
    }
    C.chunk_bz0.push_back(nbz); C.chunk_e0.push_back(C.nint); C.chunk_s0.push_back((int)C.t_s.size() / 2);
    C.allowed = 0;
    for (const auto& v : per_bz) C.allowed += (long long)v.size();
    {   // unique AO quartets (the reference's pair12 >= pair34 enumeration, pyx:1312-1331) per degeneracy case
        auto count = [&](bool ab, bool cd, bool diag) {
            double n = 0;
            for (int a = 0; a < ncA; ++a) for (int b = 0; b < ncB; ++b) {
                if (ab && b > a) continue;
                for (int c = 0; c < ncC; ++c) for (int d = 0; d < ncD; ++d) {
                    if (cd && d > c) continue;
                    if ((T.pg[La][a] ^ T.pg[Lb][b] ^ T.pg[Lc][c] ^ T.pg[Ld][d]) != 0) continue;
                    if (diag && (c * ncD + d) > (a * ncB + b)) continue;
                    n += 1;
                }
            }
            return n;
        };
        C.uniq[0] = count(false, false, false); C.uniq[1] = count(true, false, false); C.uniq[2] = count(false, true, false);
        C.uniq[3] = count(true, true, false);
        C.uniq[4] = (La == Lc && Lb == Ld) ? count(false, false, true) : 0;
        C.uniq[5] = (La == Lb && Lb == Lc && Lc == Ld) ? count(true, true, true) : 0;
    }
    const int nchunk = (int)C.chunk_bz0.size() - 1;
    for (int ch = 0; ch < nchunk; ++ch)
        sort_pairs(C.t_s, 2 * (size_t)C.chunk_s0[ch], 2 * (size_t)C.chunk_s0[ch + 1], [](const std::pair<unsigned, unsigned>& e) { return e.second >> 16; });
    // integral slots of a chunk re-ordered by descending number of (m, m') terms
    std::vector<int> slot_perm(C.nint);       // old global slot (chunk_e0 + slot) -> new slot inside its chunk
    for (int ch = 0; ch < nchunk; ++ch) {
        const int e0 = C.chunk_e0[ch], ne = C.chunk_e0[ch + 1] - e0;
        std::vector<int> order(ne);
        for (int i = 0; i < ne; ++i) order[i] = i;
        auto nterm = [&](int i) {
            const unsigned w1 = C.p4[2 * (size_t)(e0 + i) + 1];
            const int mx = (int)((w1 >> 20) & 15) - (int)((w1 >> 16) & 15) + 1, my = (int)((w1 >> 28) & 15) - (int)((w1 >> 24) & 15) + 1;
            return mx * my;
        };
#ifdef TUNA_SHELL_ASM_UNROLL
        // the assembly dispatches on the m' trip count: lanes of a warp should agree on it first, then on the number of m trips
        auto akey = [&](int i) {
            const unsigned w1 = C.p4[2 * (size_t)(e0 + i) + 1];
            const int mx = (int)((w1 >> 20) & 15) - (int)((w1 >> 16) & 15) + 1, my = (int)((w1 >> 28) & 15) - (int)((w1 >> 24) & 15) + 1;
            return my * 16 + mx;
        };
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return akey(x) > akey(y); });
#else
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return nterm(x) > nterm(y); });
#endif
        std::vector<unsigned> np4(2 * (size_t)ne);
        for (int k = 0; k < ne; ++k) {
            np4[2 * k] = C.p4[2 * (size_t)(e0 + order[k])];
            np4[2 * k + 1] = C.p4[2 * (size_t)(e0 + order[k]) + 1];
            slot_perm[e0 + order[k]] = k;
        }
        std::copy(np4.begin(), np4.end(), C.p4.begin() + 2 * (size_t)e0);
    }
    // phase 5: CSR over accumulators per chunk; accumulators ordered by descending total number of terms (same order in every chunk)
    // distinct pair-function quartets of every chunk with their (beta, gamma) indices, for the J terms
    struct PFQ { int slot, beta, gamma; };
    std::vector<std::vector<PFQ>> pfq(nchunk);
    {
        std::map<long long, int> seen;
        for (int ch = 0; ch < nchunk; ++ch)
            for (int bi = C.chunk_bz0[ch]; bi < C.chunk_bz0[ch + 1]; ++bi)
                for (const Quartet& q : per_bz[bi]) {
                    const long long key = pf_key(bi, q);
                    if (seen.count(key)) continue;
                    seen[key] = 1;
                    pfq[ch].push_back({slot_perm[C.chunk_e0[ch] + slot_of[key]], bidx[q.a * ncB + q.b], gidx[q.c * ncD + q.d]});
                }
    }
    // fill list (dense-tensor mode): every parity-allowed component quartet of the chunk with the It slot it reads
    C.chunk_f0.push_back(0);
    for (int ch = 0; ch < nchunk; ++ch) {
        for (int bi = C.chunk_bz0[ch]; with_fill && bi < C.chunk_bz0[ch + 1]; ++bi)
            for (const Quartet& q : per_bz[bi]) {
                const unsigned it = (unsigned)slot_perm[C.chunk_e0[ch] + slot_of[pf_key(bi, q)]];
                C.p6.push_back(it | (unsigned)q.a << 16 | (unsigned)q.b << 21 | (unsigned)q.c << 26);
                C.p6.push_back((unsigned)q.d);
            }
        C.chunk_f0.push_back((int)C.p6.size() / 2);
    }
    std::vector<long long> tot_terms(C.nout, 0);
    for (const auto& v : per_bz)
        for (const Quartet& q : v) {
            ++tot_terms[ob[0] + q.a * ncC + q.c]; ++tot_terms[ob[1] + q.a * ncD + q.d]; ++tot_terms[ob[2] + q.b * ncC + q.c];
            ++tot_terms[ob[3] + q.b * ncD + q.d];
        }
    for (const auto& v : pfq)
        for (const PFQ& f : v) { ++tot_terms[C.nk + f.beta]; ++tot_terms[C.nk + nbeta + f.gamma]; }
    std::vector<int> oorder(C.nout), opos(C.nout);       // new position -> old accumulator id, and the inverse
    for (int o = 0; o < C.nout; ++o) oorder[o] = o;
    std::stable_sort(oorder.begin(), oorder.end(), [&](int x, int y) { return tot_terms[x] > tot_terms[y]; });
    for (int k = 0; k < C.nout; ++k) opos[oorder[k]] = k;
    {
        std::vector<unsigned short> nomap(C.nout);
        for (int k = 0; k < C.nout; ++k) nomap[k] = C.omap[oorder[k]];
        C.omap.swap(nomap);
    }
    // J flush list: every component pair reads its pair-function accumulator
    for (int a = 0; a < ncA; ++a)
        for (int b = 0; b < ncB; ++b) C.jflush.push_back((unsigned)rc(0, a, 1, b) | (unsigned)opos[C.nk + bidx[a * ncB + b]] << 16);
    for (int c = 0; c < ncC; ++c)
        for (int d = 0; d < ncD; ++d) C.jflush.push_back((unsigned)rc(2, c, 3, d) | (unsigned)opos[C.nk + nbeta + gidx[c * ncD + d]] << 16);
    for (int ch = 0; ch < nchunk; ++ch) {
        std::vector<std::vector<unsigned>> terms(C.nout);
        for (int bi = C.chunk_bz0[ch]; bi < C.chunk_bz0[ch + 1]; ++bi)
            for (const Quartet& q : per_bz[bi]) {
                const unsigned it = (unsigned)slot_perm[C.chunk_e0[ch] + slot_of[pf_key(bi, q)]];
                terms[opos[ob[0] + q.a * ncC + q.c]].push_back(it | (unsigned)(pb[0] + q.d * ncB + q.b) << 16);   // KAC += I P[d][b]
                terms[opos[ob[1] + q.a * ncD + q.d]].push_back(it | (unsigned)(pb[1] + q.c * ncB + q.b) << 16);   // KAD += I P[c][b]
                terms[opos[ob[2] + q.b * ncC + q.c]].push_back(it | (unsigned)(pb[2] + q.d * ncA + q.a) << 16);   // KBC += I P[d][a]
                terms[opos[ob[3] + q.b * ncD + q.d]].push_back(it | (unsigned)(pb[3] + q.c * ncA + q.a) << 16);   // KBD += I P[c][a]
            }
        for (const PFQ& f : pfq[ch]) {
            terms[opos[C.nk + f.beta]].push_back((unsigned)f.slot | (unsigned)(C.nk + f.gamma) << 16);                    // Jb[beta]  += I Pg[gamma]
            terms[opos[C.nk + nbeta + f.gamma]].push_back((unsigned)f.slot | (unsigned)(C.nk + ngamma + f.beta) << 16);   // Jg[gamma] += I Pb[beta]
        }
        // transposed storage in blocks of 32 accumulators: quad t of accumulator o at ((ptr[o / 32] + t * 32 + o % 32) * 4 words;
        // every list of a block is padded to the block's longest (dummy term: zero slot It[itmax], P stage entry 0)
        C.p5off.push_back((unsigned)C.p5term.size());       // multiple of 4: 16-byte aligned uint4 loads
        unsigned run = 0;                                   // in units of four terms
        const int nblk = (C.nout + 31) / 32;
        for (int blk = 0; blk < nblk; ++blk) {
            C.p5ptr.push_back(run);
            size_t mx = 0;
            for (int o = blk * 32; o < std::min(C.nout, blk * 32 + 32); ++o) mx = std::max(mx, (terms[o].size() + 3) / 4);
            const size_t base = C.p5term.size();
            C.p5term.resize(base + mx * 128, (unsigned)C.itmax);
            for (int o = blk * 32; o < std::min(C.nout, blk * 32 + 32); ++o)
                for (size_t k = 0; k < terms[o].size(); ++k) C.p5term[base + ((k / 4) * 32 + (size_t)(o & 31)) * 4 + (k & 3)] = terms[o][k];
#ifdef TUNA_SHELL_WIDE_TERMS
            {   // same blocks, two 16-byte entries per quad: half h = (k & 3) >> 1 of quad k / 4 at uint4 index 2 * run + (2 (k / 4) + h) * 32 + lane
                const size_t wbase = 2 * (size_t)C.p5off.back() + 8 * (size_t)run;
                C.p5wide.resize(wbase + mx * 256, 0u);
                for (size_t e = wbase; e < wbase + mx * 256; e += 2) C.p5wide[e] = (unsigned)C.itmax;      // padding: zero slot, density entry 0
                for (int o = blk * 32; o < std::min(C.nout, blk * 32 + 32); ++o)
                    for (size_t k = 0; k < terms[o].size(); ++k) {
                        const size_t w = wbase + 4 * ((2 * (k / 4) + ((k & 3) >> 1)) * 32 + (size_t)(o & 31)) + 2 * (k & 1);
                        C.p5wide[w] = terms[o][k] & 0xffffu;
                        C.p5wide[w + 1] = terms[o][k] >> 16;
                    }
            }
#endif
            run += (unsigned)mx * 32;
        }
        C.p5ptr.push_back(run);
        for (int o = 0; o < C.nout; ++o) C.p5real += (long long)terms[o].size();
    }
}

// View of the tables with the given base pointers (host vectors for the CPU test build, device copies for the GPU).
template <class PtrOf>
inline ClassTablesDev class_tables_view(const ClassTablesHost& C, PtrOf ptr) {
    ClassTablesDev V;
    V.nchunk = (int)C.chunk_bz0.size() - 1; V.nout = C.nout; V.itmax = C.itmax; V.smax_rows = C.smax_rows; V.nk = C.nk;
    V.chunk_bz0 = ptr(C.chunk_bz0); V.chunk_e0 = ptr(C.chunk_e0); V.chunk_s0 = ptr(C.chunk_s0);
    V.p4 = ptr(C.p4); V.p5ptr = ptr(C.p5ptr); V.p5term = ptr(C.p5term); V.p5off = ptr(C.p5off);
    V.pmap = ptr(C.pmap); V.omap = ptr(C.omap);
    V.njst = (int)C.jst_ptr.size() - 1; V.njfl = (int)C.jflush.size();
    V.jst_ptr = ptr(C.jst_ptr); V.jst_list = ptr(C.jst_list); V.jflush = ptr(C.jflush);
    V.n_rt = (int)C.t_rt.size(); V.n_xy = (int)C.t_xy.size(); V.n_u = (int)C.t_u.size() / 2;
    V.t_rt = ptr(C.t_rt); V.t_xy = ptr(C.t_xy); V.t_u = ptr(C.t_u); V.t_s = ptr(C.t_s);
    V.p6 = ptr(C.p6); V.chunk_f0 = ptr(C.chunk_f0);
#ifdef TUNA_SHELL_WIDE_TERMS
    V.p5w = nullptr;      // set by the caller from scale_wide_terms(C.p5wide, nb)
#endif
    return V;
}

#ifdef TUNA_SHELL_WIDE_TERMS
// Device (or emulation) copy of the wide phase-5 table for a job that batches nb quartets: byte offsets = element offsets * 8 nb.
// Density words additionally carry p_bias = oP - oIt of the job's shared-memory layout, so that both operands are addressed from
// the It buffer.
inline std::vector<unsigned> scale_wide_terms(const std::vector<unsigned>& wide, int nb, int p_bias) {
    std::vector<unsigned> out(wide.size());
    for (size_t i = 0; i < wide.size(); ++i) out[i] = (wide[i] + ((i & 1) ? (unsigned)p_bias : 0u)) * 8u * (unsigned)nb;
    return out;
}
#endif

struct HostPtrOf {
    template <class V> const typename V::value_type* operator()(const V& v) const { return v.data(); }
};

}  // namespace tuna
