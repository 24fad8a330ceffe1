// pairtable.hpp — host-side construction of the AO-pair / primitive-pair table, built once per geometry.
//
// Replaces build_primitive_pair_eri / build_ao_pair_eri / fill_hermite_table_iter_eri
// (TUNA/tuna_integrals/tuna_integral.pyx:1050-1128, :961-1036) and the serial pair loop at :1300-1308.
// Differences by design: records are 26 doubles (only parity-allowed x/y Hermite entries are kept), and
// pairs are SORTED by (x/y parity class, angular class, contraction length) so that a warp of the quartet
// kernels sees uniform loop bounds and the parity test of pyx:1324-1327 becomes "same parity group".
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "eri_core.cuh"

namespace tuna {

struct HostBasis {
    int ncart = 0;
    std::vector<double> oz;          // z of the centre
    std::vector<int> lmn;            // [ncart][3]
    std::vector<int> nprim;
    std::vector<int64_t> off;
    std::vector<double> exps, ceff;  // ceff = norm[k] * coefs[k]  (pyx:1070)
};

struct PairTable {
    int64_t npair = 0;
    std::vector<int> pi, pj;             // AO indices, pi >= pj
    std::vector<int> cls;                // lx | ly << 8 | lz << 16
    std::vector<int> npp;                // primitive pairs in this AO pair
    std::vector<int64_t> ppoff;          // record offset (in records)
    std::vector<double> pp;              // records, PP_DOUBLES each
    int64_t group_begin[5] = {0, 0, 0, 0, 0};   // parity groups: (lx&1) * 2 + (ly&1)
};

// Hermite expansion coefficients E_t^{i j}, 0 <= t <= i + j, for one Cartesian direction.
// E is filled for all i <= l1, j <= l2 at stride (l1+l2+2); returns pointer semantics via index helper.
#ifdef _OPENMP
inline int host_threads() { return omp_get_max_threads(); }
inline int host_thread_id() { return omp_get_thread_num(); }
#else
inline int host_threads() { return 1; }
inline int host_thread_id() { return 0; }
#endif
constexpr size_t HERMITE_SCRATCH_DOUBLES = 4096;      // >= (l1 + 1)(l2 + 1)(l1 + l2 + 2) for every table the host builds (l1 + l2 <= 10)

inline void hermite_table(int l1, int l2, double R, double a, double b, std::vector<double>& E) {
    const int nt = l1 + l2 + 2;
    E.assign((size_t)(l1 + 1) * (l2 + 1) * nt, 0.0);
    const double p = a + b, mu = a * b / p, h = 0.5 / p;
    const double xpa = -b / p * R, xpb = a / p * R;
    auto at = [&](int i, int j) { return E.data() + ((size_t)i * (l2 + 1) + j) * nt; };
    at(0, 0)[0] = std::exp(-mu * R * R);
    for (int i = 0; i <= l1; ++i)
        for (int j = 0; j <= l2; ++j) {
            if (i == 0 && j == 0) continue;
            double* cur = at(i, j);
            const double* prev = (j == 0) ? at(i - 1, 0) : at(i, j - 1);
            const double x = (j == 0) ? xpa : xpb;
            for (int t = 0; t <= i + j; ++t) {
                double v = x * prev[t] + (t + 1) * prev[t + 1];
                if (t > 0) v += h * prev[t - 1];
                cur[t] = v;
            }
        }
}

inline void fill_prim_record(double* rec, const HostBasis& B, int i, int j, int64_t ia, int64_t jb, std::vector<double>& scratch) {
    const int* si = &B.lmn[3 * i];
    const int* sj = &B.lmn[3 * j];
    const double a = B.exps[ia], b = B.exps[jb], p = a + b;
    for (int k = 0; k < PP_DOUBLES; ++k) rec[k] = 0.0;
    rec[0] = B.ceff[ia] * B.ceff[jb];
    rec[1] = p;
    rec[2] = (a * B.oz[i] + b * B.oz[j]) / p;
    const int lx = si[0] + sj[0], ly = si[1] + sj[1], lz = si[2] + sj[2];
    hermite_table(si[0], sj[0], 0.0, a, b, scratch);
    {
        const double* row = scratch.data() + ((size_t)si[0] * (sj[0] + 1) + sj[0]) * (lx + 2);
        for (int t = lx & 1, k = 0; t <= lx; t += 2, ++k) rec[PP_EX + k] = row[t];
    }
    hermite_table(si[1], sj[1], 0.0, a, b, scratch);
    {
        const double* row = scratch.data() + ((size_t)si[1] * (sj[1] + 1) + sj[1]) * (ly + 2);
        for (int t = ly & 1, k = 0; t <= ly; t += 2, ++k) rec[PP_EY + k] = row[t];
    }
    hermite_table(si[2], sj[2], B.oz[i] - B.oz[j], a, b, scratch);
    {
        const double* row = scratch.data() + ((size_t)si[2] * (sj[2] + 1) + sj[2]) * (lz + 2);
        for (int t = 0; t <= lz; ++t) rec[PP_EZ + t] = row[t];
    }
}

inline void build_pair_table(const HostBasis& B, PairTable& T) {
    const int n = B.ncart;
    const int64_t npair = (int64_t)n * (n + 1) / 2;
    struct Key { int grp, lx, ly, lz, npp, i, j; };
    std::vector<Key> keys;
    keys.reserve(npair);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            Key k;
            k.lx = B.lmn[3 * i] + B.lmn[3 * j];
            k.ly = B.lmn[3 * i + 1] + B.lmn[3 * j + 1];
            k.lz = B.lmn[3 * i + 2] + B.lmn[3 * j + 2];
            k.grp = (k.lx & 1) * 2 + (k.ly & 1);
            k.npp = B.nprim[i] * B.nprim[j];
            k.i = i; k.j = j;
            keys.push_back(k);
        }
    std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
        if (a.grp != b.grp) return a.grp < b.grp;
        int la = a.lx + a.ly + a.lz, lb = b.lx + b.ly + b.lz;
        if (la != lb) return la < lb;
        if (a.lx != b.lx) return a.lx < b.lx;
        if (a.ly != b.ly) return a.ly < b.ly;
        if (a.lz != b.lz) return a.lz < b.lz;
        return a.npp < b.npp;
    });
    T.npair = npair;
    T.pi.resize(npair); T.pj.resize(npair); T.cls.resize(npair); T.npp.resize(npair); T.ppoff.resize(npair);
    int64_t total = 0;
    for (int g = 0; g < 5; ++g) T.group_begin[g] = 0;
    for (int64_t a = 0; a < npair; ++a) {
        const Key& k = keys[a];
        T.pi[a] = k.i; T.pj[a] = k.j;
        T.cls[a] = k.lx | (k.ly << 8) | (k.lz << 16);
        T.npp[a] = k.npp;
        T.ppoff[a] = total;
        total += k.npp;
        T.group_begin[k.grp + 1] = a + 1;
    }
    for (int g = 1; g < 5; ++g) T.group_begin[g] = std::max(T.group_begin[g], T.group_begin[g - 1]);
    T.pp.assign((size_t)total * PP_DOUBLES, 0.0);
    // per-thread scratch is allocated OUTSIDE the parallel region (an allocation failure inside it could not be reported, only terminate);
    // hermite_table stays within the reserved capacity
    std::vector<std::vector<double>> scratch_of((size_t)host_threads());
    for (auto& v : scratch_of) v.reserve(HERMITE_SCRATCH_DOUBLES);
#pragma omp parallel
    {
        std::vector<double>& scratch = scratch_of[(size_t)host_thread_id()];
#pragma omp for schedule(dynamic, 256)
        for (int64_t a = 0; a < npair; ++a) {
            const int i = T.pi[a], j = T.pj[a];
            double* rec = T.pp.data() + (size_t)T.ppoff[a] * PP_DOUBLES;
            for (int ka = 0; ka < B.nprim[i]; ++ka)
                for (int kb = 0; kb < B.nprim[j]; ++kb, rec += PP_DOUBLES)
                    fill_prim_record(rec, B, i, j, B.off[i] + ka, B.off[j] + kb, scratch);
        }
    }
}

// Boys table F_m(T_i) by the all-positive Kummer series in long double (~1e-19 relative).
inline void build_boys_table(std::vector<double>& tab) {
    tab.assign((size_t)BOYS_ROWS * BOYS_COLS, 0.0);
    for (int i = 0; i < BOYS_ROWS; ++i) {
        const long double T = (long double)i / BOYS_INV_STEP;
        const long double e = expl(-T);
        for (int m = 0; m < BOYS_COLS; ++m) {
            long double term = 1.0L / (2 * m + 1), sum = term;
            for (int k = 1; k < 2000; ++k) {
                term *= 2.0L * T / (2 * m + 2 * k + 1);
                sum += term;
                if (term < 1e-22L * sum) break;
            }
            tab[(size_t)i * BOYS_COLS + m] = (double)(e * sum);
        }
    }
}

// a(w,k) = w! / (k! (w-2k)! 2^k)
inline void build_hermite_poly_table(std::vector<double>& herm) {
    herm.assign((size_t)(2 * L_PAIR_MAX + 1) * HERM_STRIDE, 0.0);
    for (int w = 0; w <= 2 * L_PAIR_MAX; ++w) {
        // recursion on coefficient vectors: c_w[j] = c_{w-1}[j-1] (x PQz) + (w-1) c_{w-2}[j-1]; a(w,k) with k = w - j
        for (int k = 0; 2 * k <= w; ++k) {
            long double v = 1.0L;
            for (int s = 1; s <= w; ++s) v *= s;
            for (int s = 1; s <= k; ++s) v /= s;
            for (int s = 1; s <= w - 2 * k; ++s) v /= s;
            for (int s = 0; s < k; ++s) v /= 2;
            herm[(size_t)w * HERM_STRIDE + k] = (double)v;
        }
    }
}

}  // namespace tuna
