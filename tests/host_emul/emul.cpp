// tests/host_emul/emul.cpp — CPU unit-test harness for the kernel math of tuna_b200/csrc.
//
// TEST TOOL, NOT A FALLBACK: the development container has no GPU, so the `__host__ __device__` math in
// eri_core.cuh / pairtable.hpp is compiled here with g++ and checked against the oracle before GPU time
// is spent.  The package never loads this library; the product path is the CUDA library only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../tuna_b200/csrc/pairtable.hpp"

using namespace tuna;

static HostBasis make_basis(int ncart, const double* oz, const int* lmn, const int* nprim, const int64_t* off,
                            const double* exps, const double* ceff) {
    HostBasis B;
    B.ncart = ncart;
    B.oz.assign(oz, oz + ncart);
    B.lmn.assign(lmn, lmn + 3 * ncart);
    B.nprim.assign(nprim, nprim + ncart);
    B.off.assign(off, off + ncart);
    int64_t tot = 0;
    for (int i = 0; i < ncart; ++i) tot += nprim[i];
    B.exps.assign(exps, exps + tot);
    B.ceff.assign(ceff, ceff + tot);
    return B;
}

extern "C" {

double emul_boys(int M, int m, double T) {
    static std::vector<double> tab;
    if (tab.empty()) build_boys_table(tab);
    double F[BOYS_MMAX + 1];
    boys_fill(tab.data(), M, T, F);
    return F[m];
}

int emul_eri_fill(int ncart, const double* oz, const int* lmn, const int* nprim, const int64_t* off, const double* exps,
                  const double* ceff, double* out) {
    HostBasis B = make_basis(ncart, oz, lmn, nprim, off, exps, ceff);
    PairTable T;
    build_pair_table(B, T);
    std::vector<double> boys, herm;
    build_boys_table(boys);
    build_hermite_poly_table(herm);
    const int64_t n = ncart, n2 = n * n, n3 = n2 * n;
    std::memset(out, 0, sizeof(double) * n3 * n);
    for (int g = 0; g < 4; ++g) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int64_t a = T.group_begin[g]; a < T.group_begin[g + 1]; ++a)
            for (int64_t b = T.group_begin[g]; b <= a; ++b) {
                PairClass ca{T.cls[a] & 255, (T.cls[a] >> 8) & 255, T.cls[a] >> 16};
                PairClass cb{T.cls[b] & 255, (T.cls[b] >> 8) & 255, T.cls[b] >> 16};
                double v = eri_ao_quartet(T.pp.data() + T.ppoff[a] * PP_DOUBLES, T.npp[a], T.pp.data() + T.ppoff[b] * PP_DOUBLES,
                                          T.npp[b], ca, cb, boys.data(), herm.data());
                int64_t i = T.pi[a], j = T.pj[a], k = T.pi[b], l = T.pj[b];
                out[i * n3 + j * n2 + k * n + l] = v; out[k * n3 + l * n2 + i * n + j] = v;
                out[j * n3 + i * n2 + l * n + k] = v; out[l * n3 + k * n2 + j * n + i] = v;
                out[j * n3 + i * n2 + k * n + l] = v; out[l * n3 + k * n2 + i * n + j] = v;
                out[i * n3 + j * n2 + l * n + k] = v; out[k * n3 + l * n2 + j * n + i] = v;
            }
    }
    return 0;
}
}
