// pp_small.cuh — building blocks of the one-block-per-sample kernels for small grids
// (P = G*G <= 64).  Everything lives in shared memory.  [C x P] operands keep the DENSE global
// layout (row stride P), so a sample is staged with straight float4 copies whose loads are all
// issued before the first store (one block per SM has only 8 warps: memory-level parallelism
// has to come from each thread).  P x P matrices use a row stride of PS (float4 rows).
#pragma once
#include "pp_common.cuh"

namespace pp {

constexpr int SM_THREADS = 256;
constexpr int PMAX = 64;
constexpr int PS = PMAX + 4;   // row stride of the P x P matrices (multiple of 4: float4 rows)
constexpr int SLACK = 8;       // floats of slack after each [C x P] buffer (4-wide tile reads overrun a row end)

// smem <- global, n floats.  Vector path when n % 4 == 0 and src is 16-byte aligned.
template <int NT = SM_THREADS>
__device__ __forceinline__ void stage_dense(float* __restrict__ dst, const float* __restrict__ src, int n) {
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const int n4 = n >> 2;
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int base = 0; base < n4; base += NT * 8) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int e = base + u * NT + threadIdx.x;
                if (e < n4) v[u] = __ldg(s4 + e);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                int e = base + u * NT + threadIdx.x;
                if (e < n4) d4[e] = v[u];
            }
        }
    } else {
        for (int e = threadIdx.x; e < n; e += NT) dst[e] = __ldg(src + e);
    }
}

// column reduction over C of f(row c, column i); result in res[i], i < P.  All NT threads call.
// red: [NT/64][PMAX] floats.
template <int NT = SM_THREADS, class F>
__device__ __forceinline__ void col_reduce(int C, int P, float* red, float* res /*[PMAX]*/, F f) {
    constexpr int NPART = NT / 64;
    const int i = threadIdx.x & 63, part = threadIdx.x >> 6;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (i < P) {
        int c = part;
        for (; c + 3 * NPART < C; c += 4 * NPART) {  // 4 independent accumulators: loads of f() overlap
            s0 += f(c, i);
            s1 += f(c + NPART, i);
            s2 += f(c + 2 * NPART, i);
            s3 += f(c + 3 * NPART, i);
        }
        for (; c < C; c += NPART) s0 += f(c, i);
    }
    if (i < PMAX) red[part * PMAX + i] = (s0 + s1) + (s2 + s3);  // a part is 64 threads wide, a row of `red` PMAX floats
    __syncthreads();
    if (threadIdx.x < PMAX) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < NPART; q++) t += red[q * PMAX + threadIdx.x];
        res[threadIdx.x] = t;
    }
    __syncthreads();
}

// M[i][j] = sum_c a[c][i] * b[c][j] for i,j < P (4x4 register tile per thread, 16x16 threads);
// ep(i, j, value) consumes the result.  Rows are dense (stride P): the 4-wide reads of the last
// tile run past the row end into the next row (finite data, results discarded).
template <class EP>
__device__ __forceinline__ void gram_tile(const float* __restrict__ a, const float* __restrict__ b, int C, int P, EP ep) {
    const int ti = threadIdx.x >> 4, tj = threadIdx.x & 15;
    const int i0 = ti * 4, j0 = tj * 4;
    if (i0 >= P || j0 >= P) return;
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0f;
#pragma unroll 4
    for (int c = 0; c < C; c++) {
        const float* ar = a + c * P + i0;
        const float* br = b + c * P + j0;
        const float av[4] = {ar[0], ar[1], ar[2], ar[3]}, bv[4] = {br[0], br[1], br[2], br[3]};
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++)
            if (i0 + u < P && j0 + v < P) ep(i0 + u, j0 + v, acc[u][v]);
}

// Same contraction with the C range split over NT/256 groups of 256 threads (each a 16x16 grid of 4x4
// tiles): groups > 0 park their partial tiles in `part` ([NT/256 - 1][PMAX*PS] floats), group 0 adds
// them in a fixed order and runs the epilogue.  All NT threads call (contains a __syncthreads).
template <int NT, class EP>
__device__ __forceinline__ void gram_tile_split(const float* __restrict__ a, const float* __restrict__ b, int C, int P,
                                                float* __restrict__ part, EP ep) {
    constexpr int NG = NT / 256;
    const int grp = threadIdx.x >> 8, t = threadIdx.x & 255;
    const int ti = t >> 4, tj = t & 15;
    const int i0 = ti * 4, j0 = tj * 4;
    const bool active = i0 < P && j0 < P;
    float acc[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
        for (int v = 0; v < 4; v++) acc[u][v] = 0.0f;
    if (active) {
        const int cb = (C * grp) / NG, ce = (C * (grp + 1)) / NG;
#pragma unroll 4
        for (int c = cb; c < ce; c++) {
            const float* ar = a + c * P + i0;
            const float* br = b + c * P + j0;
            const float av[4] = {ar[0], ar[1], ar[2], ar[3]}, bv[4] = {br[0], br[1], br[2], br[3]};
#pragma unroll
            for (int u = 0; u < 4; u++)
#pragma unroll
                for (int v = 0; v < 4; v++) acc[u][v] = fmaf(av[u], bv[v], acc[u][v]);
        }
        if (grp > 0) {
            float* pt = part + (grp - 1) * PMAX * PS;
#pragma unroll
            for (int u = 0; u < 4; u++)
                *reinterpret_cast<float4*>(pt + (i0 + u) * PS + j0) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        }
    }
    __syncthreads();
    if (active && grp == 0) {
#pragma unroll
        for (int g = 1; g < NG; g++) {
            const float* pt = part + (g - 1) * PMAX * PS;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 v = *reinterpret_cast<const float4*>(pt + (i0 + u) * PS + j0);
                acc[u][0] += v.x; acc[u][1] += v.y; acc[u][2] += v.z; acc[u][3] += v.w;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int v = 0; v < 4; v++)
                if (i0 + u < P && j0 + v < P) ep(i0 + u, j0 + v, acc[u][v]);
    }
}

// dst[c][i] = sum_j src[c][j] * M[j][i]   (thread = channel; M rows contiguous, stride PS).
// All NP*16 output columns of a row are accumulated in registers before anything is written,
// so dst may alias src (thread-private rows).  Row stride P is odd for odd grids: the per-thread
// row reads are bank-conflict free; M rows are warp broadcasts.
template <int NP>
__device__ __forceinline__ void row_times_mat_np(const float* src, const float* __restrict__ M, int C, int P, float* dst) {
    for (int c = threadIdx.x; c < C; c += SM_THREADS) {
        const float* s = src + c * P;
        float acc[NP][16];
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int u = 0; u < 16; u++) acc[p][u] = 0.0f;
        for (int j = 0; j < P; j++) {
            const float sj = s[j];
            const float4* m4 = reinterpret_cast<const float4*>(M + j * PS);
#pragma unroll
            for (int p = 0; p < NP; p++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float4 m = m4[4 * p + q];
                    acc[p][4 * q + 0] = fmaf(sj, m.x, acc[p][4 * q + 0]);
                    acc[p][4 * q + 1] = fmaf(sj, m.y, acc[p][4 * q + 1]);
                    acc[p][4 * q + 2] = fmaf(sj, m.z, acc[p][4 * q + 2]);
                    acc[p][4 * q + 3] = fmaf(sj, m.w, acc[p][4 * q + 3]);
                }
        }
#pragma unroll
        for (int p = 0; p < NP; p++)
#pragma unroll
            for (int u = 0; u < 16; u++)
                if (16 * p + u < P) dst[c * P + 16 * p + u] = acc[p][u];
    }
}
__device__ __forceinline__ void row_times_mat(const float* src, const float* __restrict__ M, int C, int P, float* dst) {
    if (P <= 16) row_times_mat_np<1>(src, M, C, P, dst);
    else if (P <= 32) row_times_mat_np<2>(src, M, C, P, dst);
    else if (P <= 48) row_times_mat_np<3>(src, M, C, P, dst);
    else row_times_mat_np<4>(src, M, C, P, dst);
}

// row_times_mat with the 16-column blocks of a row split over NT/C thread groups (thread = channel x
// column group).  Accumulates in registers, synchronises, then writes: dst may alias src.
// All NT threads call; requires NT % C == 0.
template <int NT>
__device__ __forceinline__ void row_times_mat_split(const float* src, const float* __restrict__ M, int C, int P, float* dst) {
    const int ngrp = NT / C;                    // column groups (2 for C = 256, NT = 512)
    const int c = threadIdx.x % C, grp = threadIdx.x / C;
    const int nblk = (P + 15) >> 4;             // 16-column blocks in a row (<= 4)
    const int per = (nblk + ngrp - 1) / ngrp;   // blocks per group (<= 2 for ngrp >= 2)
    const int b0 = grp * per, b1 = min(nblk, b0 + per);
    float acc[2][16];
#pragma unroll
    for (int p = 0; p < 2; p++)
#pragma unroll
        for (int u = 0; u < 16; u++) acc[p][u] = 0.0f;
    if (grp < ngrp && b0 < b1) {
        const float* s = src + c * P;
        for (int j = 0; j < P; j++) {
            const float sj = s[j];
            const float4* m4 = reinterpret_cast<const float4*>(M + j * PS) + 4 * b0;
#pragma unroll
            for (int p = 0; p < 2; p++)
                if (b0 + p < b1) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float4 m = m4[4 * p + q];
                        acc[p][4 * q + 0] = fmaf(sj, m.x, acc[p][4 * q + 0]);
                        acc[p][4 * q + 1] = fmaf(sj, m.y, acc[p][4 * q + 1]);
                        acc[p][4 * q + 2] = fmaf(sj, m.z, acc[p][4 * q + 2]);
                        acc[p][4 * q + 3] = fmaf(sj, m.w, acc[p][4 * q + 3]);
                    }
                }
        }
    }
    __syncthreads();  // every reader of src is done before anyone overwrites it
    if (grp < ngrp) {
#pragma unroll
        for (int p = 0; p < 2; p++)
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const int col = 16 * (b0 + p) + u;
                if (b0 + p < b1 && col < P) dst[c * P + col] = acc[p][u];
            }
    }
}

// out[e] = fn(e, c, i) for every element of a dense [C x P] tile (e = c*P + i), no divisions:
// each thread walks its strided elements keeping (c, i) incrementally.
template <int NT = SM_THREADS, class F>
__device__ __forceinline__ void for_each_ci(int C, int P, F fn) {
    int e = threadIdx.x, c = e / P, i = e - c * P;
    const int dc = NT / P, di = NT - dc * P;
    for (; e < C * P; e += NT) {
        fn(e, c, i);
        c += dc;
        i += di;
        if (i >= P) { i -= P; c += 1; }
    }
}

}  // namespace pp
