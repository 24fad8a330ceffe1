// mo_transform.cuh — AO -> MO / spin-orbital four-index transformation of the dense ERI tensor.
//
// Reference: TUNA/tuna_ci.py:204-255 (transform_ERI_AO_to_MO) and :143-193 (transform_ERI_AO_to_SO): four NumPy einsums,
//   "mknl,ls->mnks" (C_1), "mnks,kr->mnrs" (C_2), "mnrs,nq->mqrs" (C_1), "mqrs,mp->prqs" (MO) / "->pqrs" (SO) (C_2),
// each an N^5 contraction of one index with the MO coefficient matrix (OpenBLAS on the host).
//
// Here every step contracts the LAST (contiguous) axis of the current tensor and writes the new index in FRONT:
//   Out[s][X] = sum_l C[l][s] * T[X][l]        X = the three leading indices flattened
// Starting from (m k n l) the four steps give (s m k n) -> (q s m k) -> (r q s m) -> (p r q s): exactly the reference's
// interleaved chemists' layout "prqs"; the spin-orbital layout "pqrs" swaps the two middle indices in the last step's
// store.  Each step is a tall-skinny FP64 GEMM (rows n^3, inner n, columns n_mo) that reads and writes the tensor once:
// 2 K M NX flops against 8 NX (K + M) bytes — for K = M = 60 about 15 flop/B, above the FP64 ridge (~5.6 flop/B at
// 37 TFLOP/s over 6.55 TB/s), so the step is FP64-pipe bound: 128 x 64 outputs per CTA on the FP64 tensor cores
// (mma.sync m8n8k4 f64, SASS DMMA), operands staged through shared memory.  The round-1 kernel (8 x 4 FMA register tile) was bound
// by LSU->register bandwidth (ncu, nbf 110: FP64 pipe 53 % active, shared-memory data pipe 74 %); the MMA fragments need 5x fewer
// operand bytes per flop.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>

namespace tuna {

constexpr int MO_TX = 128;      // X rows per CTA
constexpr int MO_TS = 64;       // output columns (s) per CTA
constexpr int MO_KC = 16;       // contraction chunk
constexpr int MO_TXP = MO_TX + 2;

// The step on the FP64 tensor cores (measured on the B200: 16.2 -> 21.2 TFLOP/s at nbf 60 -> 110 against 13.5 -> 16.3 for the 8 x 4 FMA
// tile it replaced, profiles/r02a_mo_dmma.json).  swap != 0: X = (a, b, c) with dims (d1, d2, d3) is stored at [s][b][a][c] (spin-orbital
// layout, last step only).  D (8 x 8) += A (8 x 4) B (4 x 8)
// with mma.sync.aligned.m8n8k4.row.col.f64: rows = X, columns = s, inner = l.  Fragment layout (PTX ISA, m8n8k4 .f64), with
// g = lane >> 2 and q = lane & 3:  A: a = A[g][q];  B: b = B[q][g];  C/D: c[i] = C[g][2 q + i].
// CTA tile 128 (X) x 64 (s) as in k_axis_gemm, 8 warps as 4 (X) x 2 (s), 32 x 32 outputs per warp = 4 x 4 MMA tiles: per 4-wide
// k step a lane loads 4 + 4 operand doubles for 16 MMAs (8192 flops) — 5x fewer operand bytes per flop than the 8 x 4 FMA tile.
// Row strides of 132 / 68 doubles (== 8 words mod 32) make the fragment loads of a half-warp hit 32 distinct banks.
constexpr int MO_TXQ = MO_TX + 4;
constexpr int MO_TSQ = MO_TS + 4;

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 2) k_axis_gemm(const double* __restrict__ T, const double* __restrict__ C, double* __restrict__ Out,
                                                           long long NX, int K, int M, int swap, int d1, int d2, int d3) {
    __shared__ __align__(16) double Ts[MO_KC][MO_TXQ];      // [l][x]
    __shared__ __align__(16) double Cs[MO_KC][MO_TSQ];      // [l][s]
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int wx = (warp & 3) * 32, ws = (warp >> 2) * 32;
    const int ntile_s = (M + MO_TS - 1) / MO_TS;
    const long long X0 = (long long)(blockIdx.x / ntile_s) * MO_TX;
    const int s0 = (int)(blockIdx.x % ntile_s) * MO_TS;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    const int lc = t & 15, lr = t >> 4;
    for (int l0 = 0; l0 < K; l0 += MO_KC) {
        __syncthreads();
#pragma unroll
        for (int pass = 0; pass < MO_TX / 16; ++pass) {
            const int r = pass * 16 + lr;
            const long long X = X0 + r;
            const int l = l0 + lc;
            Ts[lc][r] = (X < NX && l < K) ? T[X * K + l] : 0.0;
        }
#pragma unroll
        for (int pass = 0; pass < (MO_KC * MO_TS) / 256; ++pass) {
            const int e = pass * 256 + t, c = e / MO_TS, s = e % MO_TS;
            const int l = l0 + c;
            Cs[c][s] = (l < K && s0 + s < M) ? C[(size_t)l * M + s0 + s] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < MO_KC; kk += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = Ts[kk + q][wx + 8 * i + g];      // A[row = X][col = l]
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Cs[kk + q][ws + 8 * j + g];      // B[row = l][col = s]
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long X = X0 + wx + 8 * i + g;
        if (X >= NX) continue;
        size_t xoff = (size_t)X, sstride = (size_t)NX;
        if (swap) {       // X = (a, b, c) stored at [s][b][a][c]
            const int c3 = (int)(X % d3);
            const long long ab = X / d3;
            const int b2 = (int)(ab % d2), a1 = (int)(ab / d2);
            xoff = ((size_t)b2 * d1 + a1) * d3 + c3;
            sstride = (size_t)d1 * d2 * d3;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int s = s0 + ws + 8 * j + 2 * q + h;
                if (s < M) Out[(size_t)s * sstride + xoff] = acc[i][j][h];
            }
    }
}

// One step on `stream`: Out[M][NX] (or the swapped layout) from T[NX][K] and C[K][M].
inline cudaError_t axis_gemm(cudaStream_t stream, const double* T, const double* C, double* Out, long long NX, int K, int M, int swap, int d1,
                             int d2, int d3) {
    const long long gx = (NX + MO_TX - 1) / MO_TX;
    const int gy = (M + MO_TS - 1) / MO_TS;
    if (gx <= 0 || gy <= 0) return cudaSuccess;
    if (gx * gy > 2147483647LL) return cudaErrorInvalidConfiguration;
    k_axis_gemm<<<(unsigned)(gx * gy), 256, 0, stream>>>(T, C, Out, NX, K, M, swap, d1, d2, d3);
    return cudaGetLastError();
}

}  // namespace tuna
