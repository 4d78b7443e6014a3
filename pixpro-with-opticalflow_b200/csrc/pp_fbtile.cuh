// pp_fbtile.cuh — forward-backward consistency mask (contrast/util.py:253-297) with the gathered
// field staged in shared memory by TMA.  Included by pp_flow.cu (needs its FB helpers).
//
// Why: the gather-from-global kernels (fbmask4*_kernel) are bound by L1 data-pipe wavefronts, not
// by DRAM or issue slots (profiles/r01_n_flow_gather_ncu_summary.txt: 30 wavefronts per warp-pixel
// for 11 memory instructions — every one of the 8 tap loads of a warp touches ~3.3 cache lines,
// and the 8 loads touch the same lines again and again).  Here each TW x TH tile's gather
// footprint is brought into shared memory ONCE by a single TMA box copy and the 8 taps per pixel
// become LDS with immediate offsets.  BW is a multiple of 32 floats so that the lanes of a warp
// (consecutive columns, a few rows apart) fall into distinct banks.
//
// Structure: one CTA = one tile, 8 warps, 4 CTAs per SM (their different phases hide the
// load -> TMA -> compute dependency; a persistent producer/consumer ring was measured slower: its
// iterator / stage bookkeeping cost more issue slots than the gather it replaced).
//  * warp 0 first evaluates the warped position (same arithmetic as below) on an 8x4 lattice of
//    the tile, warp-reduces the bounding box, centres the BW x BH box on it (x origin rounded down
//    to 4 floats: TMA needs a 16-byte aligned start, measured with profiles/mb/tma_probe.cu) and
//    issues one cp.async.bulk.tensor.3d for both channels.  Rows / columns outside the frame are
//    zero-filled by TMA — exactly the value grid_sample's zero padding gives a missing tap, so
//    frame-edge pixels need no special case.
//  * every thread then handles TW/32 columns x TH/8 rows of pixels as packed fp32 pairs of two
//    rows (see fbmask4p_kernel).  A pixel whose 2x2 footprint is not inside the staged box (flow
//    rougher than the lattice predicts) reads its taps from global memory instead, in line, so the
//    result never depends on the prediction and rough fields degrade gracefully towards the
//    gather kernel's speed.
#pragma once
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "pp_common.cuh"
#include "pp_tc.cuh"

namespace pp {
namespace fbt {

struct Args {
    const float* flow[2];  // [B,2,H,W] each
    uint8_t* mask[2];      // [B,H,W]
    int H, W, ndir;
    int B, pf_samples;  // samples; L2 prefetch distance in samples (0 = off)
    float half_w, half_h, a1, a2;
    Div<DM_FAST> dw, dh;    // / (W-1), / (H-1)
    Div<DM_FAST> dw2, dh2;  // / ((W-1)/2), / ((H-1)/2)
};

__device__ unsigned long long g_redo_pixels;  // pixels that took their taps from global memory
__device__ unsigned int g_wait_timeouts;      // mbarrier waits that gave up (pipeline bug: results invalid)

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a broken pipeline is counted and the kernel runs on (the host reports it as an
// error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); spin++) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(tc::smem_u32(bar)), "r"(parity), "r"(2000u)  // suspend-time hint (ns): sleep, do not poll
            : "memory");
        if (done) return;
    }
    atomicAdd(&g_wait_timeouts, 1u);
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(tc::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(tc::smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ uint8_t* byte_ptr_at(uint8_t* base, int off) {  // base + off as one IMAD.WIDE
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"((unsigned long long)base));
    return reinterpret_cast<uint8_t*>(r);
}
// L2 prefetch of one box (no shared-memory destination, no completion to wait for)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

// The FB test of one packed pixel pair (two rows 8 apart of one column), shared by fbbox_kernel and fbbox_up_kernel: own flow
// (fxs, fys) -> normalised position (util.py:264-276) -> the 2x2 taps of the gather source from the staged box (in line from
// global memory `g` when the footprint is outside the box) -> cycle test (util.py:279-296) -> the two mask bytes.
template <int BW, int BH>
__device__ __forceinline__ void fb_pair(const float (&fxs)[2], const float (&fys)[2], const F2 xn, const F2 yn, const Div2& dW,
                                        const Div2& dH, const F2 one2, const F2 hw2, const F2 hh2, const F2 a1_2, const F2 a2_2,
                                        const float* sp, const int2 o, const float* g, const int W, const int H, const int HW,
                                        uint8_t* m0, uint8_t* m1, int& nglobal) {
    const F2 fnx = dW(pk(fxs[0], fxs[1])), fny = dH(pk(fys[0], fys[1]));                           // :264
    const F2 c1x = add2(xn, fnx), c1y = add2(yn, fny);                                             // :275
    // (c1 + 1) * half: >= 0 for every in-frame pixel, so the non-contractable product form is exact
    const F2 ix = mul2_nc(add2(c1x, one2), hw2), iy = mul2_nc(add2(c1y, one2), hh2);
    float c1xs[2], c1ys[2], ixs[2], iys[2];
    unpk(c1x, c1xs[0], c1xs[1]); unpk(c1y, c1ys[0], c1ys[1]);
    unpk(ix, ixs[0], ixs[1]); unpk(iy, iys[0], iys[1]);
    bool inb[2], outside[2];
    int gofs[2];
    float t[2][8], xws[2], yws[2];
#pragma unroll
    for (int p = 0; p < 2; p++) {
        inb[p] = (fabsf(c1xs[p]) < 1.0f) && (fabsf(c1ys[p]) < 1.0f);                               // :276
        const int x0 = __float2int_rd(ixs[p]), y0 = __float2int_rd(iys[p]);
        xws[p] = __int2float_rn(x0);
        yws[p] = __int2float_rn(y0);
        const unsigned dx = (unsigned)(x0 - o.x), dy = (unsigned)(y0 - o.y);
        outside[p] = inb[p] && (dx > (unsigned)(BW - 2) || dy > (unsigned)(BH - 2));
        gofs[p] = (y0 << 16) | (x0 & 0xffff);  // only read for in-frame pixels (launcher: W, H < 32768)
        // clamp: a pixel outside the box or the frame still addresses the staged box
        const float* q = sp + min(dy, (unsigned)(BH - 2)) * BW + min(dx, (unsigned)(BW - 2));
        t[p][0] = q[0]; t[p][1] = q[1]; t[p][2] = q[BW]; t[p][3] = q[BW + 1];
        t[p][4] = q[BW * BH]; t[p][5] = q[BW * BH + 1]; t[p][6] = q[BW * BH + BW]; t[p][7] = q[BW * BH + BW + 1];
    }
    if (outside[0] || outside[1]) {
        // footprint outside the staged box (rare): the same taps straight from global memory,
        // zero where grid_sample pads (an in-frame pixel can only miss column W or row H)
#pragma unroll
        for (int p = 0; p < 2; p++)
            if (outside[p]) {
                const int x0 = gofs[p] & 0xffff, y0 = gofs[p] >> 16;
                const float* q = ptr_at(g, y0 * W + x0);
                const bool xin = x0 < W - 1, yin = y0 < H - 1;
                t[p][0] = __ldg(q); t[p][4] = __ldg(ptr_at(q, HW));
                t[p][1] = xin ? __ldg(q + 1) : 0.0f; t[p][5] = xin ? __ldg(ptr_at(q, HW) + 1) : 0.0f;
                t[p][2] = yin ? __ldg(ptr_at(q, W)) : 0.0f; t[p][6] = yin ? __ldg(ptr_at(q, HW + W)) : 0.0f;
                t[p][3] = (xin && yin) ? __ldg(ptr_at(q, W) + 1) : 0.0f; t[p][7] = (xin && yin) ? __ldg(ptr_at(q, HW + W) + 1) : 0.0f;
                nglobal++;
            }
    }
    const F2 wx = sub2(ix, pk(xws[0], xws[1])), wy = sub2(iy, pk(yws[0], yws[1]));
    const F2 e = sub2(one2, wx), s_ = sub2(one2, wy);
    const F2 nw = mul2(s_, e), ne = mul2(s_, wx), sw = mul2(wy, e), se = mul2(wy, wx);
    const F2 bx = combine4_2(dW(pk(t[0][0], t[1][0])), dW(pk(t[0][1], t[1][1])), dW(pk(t[0][2], t[1][2])),
                             dW(pk(t[0][3], t[1][3])), nw, ne, sw, se);
    const F2 by = combine4_2(dH(pk(t[0][4], t[1][4])), dH(pk(t[0][5], t[1][5])), dH(pk(t[0][6], t[1][6])),
                             dH(pk(t[0][7], t[1][7])), nw, ne, sw, se);
    const F2 cyx = add2(fnx, bx), cyy = add2(fny, by);                                             // :279
    // squares and alpha_1 * sum are never -0 (alpha_1 >= 0 is checked by the launcher)
    const F2 cyc2 = add2(mul2_nc(cyx, cyx), mul2_nc(cyy, cyy));                                    // :293
    const F2 f2 = add2(mul2_nc(fnx, fnx), mul2_nc(fny, fny)), b2 = add2(mul2_nc(bx, bx), mul2_nc(by, by));
    const F2 eps = add2(mul2_nc(a1_2, add2(f2, b2)), a2_2);                                        // :294
    float ds[2];
    unpk(sub2(cyc2, eps), ds[0], ds[1]);
    *m0 = (inb[0] && (ds[0] <= 0.0f)) ? 1 : 0;                                                     // :296
    *m1 = (inb[1] && (ds[1] <= 0.0f)) ? 1 : 0;
}

// grid = (W / TW, H / TH, planes); block = 256.  Requires W % TW == 0 and H % TH == 0.
// WC/HC > 0: the frame size is a compile-time constant (the 1280x720 frames of the published BDD100K runs): every
// derived constant and row offset folds into instruction immediates (no per-pixel constant-bank loads).
// Five value-preserving instruction-diet switches of this kernel (immediate-offset addressing off one per-thread pointer; floor by
// a round-down add of 1.5*2^23 on packed pairs instead of F2I.FLOOR + I2FP; lazily re-derived tap origin on the rare global path;
// a rolled row loop; opaque row pointers) were built, verified bit-exact and measured within +-1.5 % of this plain form
// (profiles/r02_a_fb_variants.txt, profiles/r02_g_fb_variants.txt: 355-364 us at B=64) — the L1 data pipe is as full as the issue
// port, so shaving glue instructions does not move it — and were removed again.
template <int TW, int TH, int BW, int BH, int MINB, int WC, int HC>
__global__ void __launch_bounds__(256, MINB) fbbox_kernel(const __grid_constant__ CUtensorMap tm0,
                                                          const __grid_constant__ CUtensorMap tm1,
                                                          const __grid_constant__ CUtensorMap tp0,
                                                          const __grid_constant__ CUtensorMap tp1, Args a) {
    constexpr int DM = DM_FAST;
    constexpr int NX = TW / 32;  // columns per thread
    constexpr int NR = TH / 8;   // rows per thread (row k -> Y0 + 8 k)
    static_assert(TW % 32 == 0 && TH % 16 == 0, "tile shape");
    constexpr uint32_t BOX_BYTES = 2 * BW * BH * 4;
    extern __shared__ __align__(1024) uint8_t fbt_smem[];
    __shared__ uint64_t full;
    __shared__ int2 origin;
    __shared__ int nglobal_cta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {  // fold every derived constant
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw2 = const_div<DM>((float)(WC - 1) / 2.0f);
        a.dh2 = const_div<DM>((float)(HC - 1) / 2.0f);
        a.dw = const_div<DM>((float)(WC - 1));
        a.dh = const_div<DM>((float)(HC - 1));
    }
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const float* f = (dir ? a.flow[1] : a.flow[0]) + (int64_t)b * 2 * HW;
    if (threadIdx.x == 0) {
        nglobal_cta = 0;
        tc::mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32 && dir == 0 && b + a.pf_samples < a.B) {
        // Pull the same tile of both fields of a LATER sample into L2 (exact TW x TH tiles, maps tp*):
        // the CTAs that will work there ~one wave from now then find their lattice samples, their own
        // flow and most of their box in L2, which shortens the exposed load -> TMA -> compute chain.
        tma_prefetch_3d(&tp0, tx0, ty0, (b + a.pf_samples) * 2);
        tma_prefetch_3d(&tp1, tx0, ty0, (b + a.pf_samples) * 2);
    }
    if (warp == 0) {  // place the box and start the copy
        const int X = tx0 + ((lane & 7) * (TW - 1)) / 7, Y = ty0 + ((lane >> 3) * (TH - 1)) / 3;
        const float* p = ptr_at(f, Y * W + X);
        const float fnx = norm_flow_h(__ldg(p), a.dw2), fny = norm_flow_h(__ldg(ptr_at(p, HW)), a.dh2);
        const float c1x = add(norm_coord_h((float)X, a.dw2), fnx), c1y = add(norm_coord_h((float)Y, a.dh2), fny);
        const bool inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);
        const int x0 = __float2int_rd(mul(add(c1x, 1.0f), a.half_w)), y0 = __float2int_rd(mul(add(c1y, 1.0f), a.half_h));
        const int mnx = __reduce_min_sync(0xffffffffu, inb ? x0 : INT_MAX);
        const int mxx = __reduce_max_sync(0xffffffffu, inb ? x0 : INT_MIN);
        const int mny = __reduce_min_sync(0xffffffffu, inb ? y0 : INT_MAX);
        const int mxy = __reduce_max_sync(0xffffffffu, inb ? y0 : INT_MIN);
        if (lane == 0) {
            int ox, oy;
            if (mnx == INT_MAX) {  // no sample lands in the frame: any box will do
                ox = tx0 - (BW - TW) / 2;
                oy = ty0 - (BH - TH) / 2;
            } else {  // centre the box on the footprint [mn, mx + 1]
                ox = mnx - (BW - (mxx + 2 - mnx)) / 2;
                oy = mny - (BH - (mxy + 2 - mny)) / 2;
            }
            ox &= ~3;  // TMA: the box must start on a 16-byte boundary of the innermost dimension
            origin = make_int2(ox, oy);
            mbar_arrive_expect_tx(&full, BOX_BYTES);
            tma_load_3d(fbt_smem, dir ? &tm0 : &tm1, ox, oy, b * 2, &full);
        }
        __syncwarp();
    }
    const Div2 dW = make_div2(a.dw2), dH = make_div2(a.dh2);
    const F2 one2 = pk1(1.0f), hw2 = pk1(a.half_w), hh2 = pk1(a.half_h), a1_2 = pk1(a.a1), a2_2 = pk1(a.a2);
    const int X0 = tx0 + lane, Y0 = ty0 + warp;
    const float* fp = ptr_at(f, Y0 * W + X0);                  // own flow, x channel, row Y0
    uint8_t* mp = (dir ? a.mask[1] : a.mask[0]) + (int64_t)b * HW + Y0 * W + X0;
    const int rstep = 8 * W;                                    // one thread-row (8 image rows)
    F2 xn2[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) xn2[c] = pk1(norm_coord_h((float)(X0 + 32 * c), a.dw2));
    float nfx[NX][2], nfy[NX][2];  // own flow of the next row pair (register prefetch)
    auto prefetch = [&](int kp) {
#pragma unroll
        for (int c = 0; c < NX; c++)
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const float* q = ptr_at(fp, (kp + p) * rstep + 32 * c);
                nfx[c][p] = __ldg(q);
                nfy[c][p] = __ldg(ptr_at(q, HW));
            }
    };
    prefetch(0);
    mbar_wait_bounded(&full, 0);
    const int2 o = origin;
    const float* sp = reinterpret_cast<const float*>(fbt_smem);
    int nglobal = 0;  // pixels of this thread whose footprint was outside the staged box
    const float* g = (dir ? a.flow[0] : a.flow[1]) + (int64_t)b * 2 * HW;
#pragma unroll(NR / 2)
    for (int kp = 0; kp < NR; kp += 2) {
        float fxs[NX][2], fys[NX][2];
#pragma unroll
        for (int c = 0; c < NX; c++)
#pragma unroll
            for (int p = 0; p < 2; p++) { fxs[c][p] = nfx[c][p]; fys[c][p] = nfy[c][p]; }
        if (kp + 2 < NR) prefetch(kp + 2);
        const F2 yn = sub2(dH(pk((float)(Y0 + 8 * kp), (float)(Y0 + 8 * kp + 8))), one2);          // :271
#pragma unroll
        for (int c = 0; c < NX; c++) {
            fb_pair<BW, BH>(fxs[c], fys[c], xn2[c], yn, dW, dH, one2, hw2, hh2, a1_2, a2_2, sp, o, g, W, H, HW,
                            byte_ptr_at(mp, kp * rstep + 32 * c), byte_ptr_at(mp, (kp + 1) * rstep + 32 * c), nglobal);
        }
    }
    // diagnostics counter: one global atomic per CTA at most (per-thread atomics on one address
    // serialise in L2 and cost milliseconds on rough fields)
    const int wsum = __reduce_add_sync(0xffffffffu, nglobal);
    if (lane == 0 && wsum) atomicAdd(&nglobal_cta, wsum);
    __syncthreads();
    if (threadIdx.x == 0 && nglobal_cta) atomicAdd(&g_redo_pixels, (unsigned long long)nglobal_cta);
}

// ---------------------------------------------------------------------------------------------
// fbbox_up_kernel — the FB mask of ONE direction with that direction's own composite COMPUTED in the kernel instead of
// loaded (n == 1 with flow_up, the published n_frames = 2 setting: the composite is the x8-up-sampled low-res link,
// contrast/flow/utils/utils.py:87-89).  pp_flow_stage then runs
//     upchain1 (forward composite only)  ->  fbbox_up<WRITE> (backward: computes + writes its composite, gathers the forward
//     one, writes mask_bwd)  ->  the forward mask (fbbox_up without WRITE, or fbbox_kernel on one direction)
// instead of upchain1 (both composites) -> fbbox (both masks): the HBM-write-bound up-sampling of one direction (77 us at
// B = 64) moves into a kernel that is bound by instruction issue and has DRAM bandwidth to spare.
// Own flow, bit-identical to upchain1_kernel / ATen's upsample_bilinear2d (see pp_chainup.cuh, steps C and D): the tile's
// 64 columns x <= 8 low-res rows are interpolated HORIZONTALLY once per CTA into shared memory, pre-scaled by 8 (exact),
//     T8(r, X) = fma(l0x, 8 L[r][i0x], l1x * 8 L[r][i1x])     (both channels as one packed pair),
// and every pixel takes U(Y, X) = fma(l0y, T8(i0y, X), l1y * T8(i1y, X)) with its row's taps from a 48-entry table.
// The table fill overlaps the flight of the TMA box.  Everything after the own flow is fbbox_kernel's code.
// Shared memory: box 96 x 64 x 2 fp32 (8 rows fewer than fbbox_kernel's: measured free, profiles/r02_ab_fb_boxrows.txt:
// 358.8 vs 361.3 us, 0.66 % of the pixels on the in-line global path) | T8 [8][TW] pairs | row taps [TH].
// (Keeping a thread's table rows in registers across its rows — they are 8 apart, so the south row of one is the north row
// of the next: 7 instead of 12 table loads per column — was built and measured: no change, 207 vs 203 us; removed.)
struct UpArgs {
    const float* lo;     // own direction's low-res link of sample b at lo + b * lo_stride, [2,h,w]
    int64_t lo_stride;
    float* own_out;      // WRITE: [B,2,H,W], receives the own composite
    const float* other;  // the opposite composite [B,2,H,W], complete in memory: the gather source
    uint8_t* mask;       // [B,H,W]
    int H, W, h, w, B, pf_samples;
    float rh, rw, half_w, half_h, a1, a2;
    Div<DM_FAST> dw2, dh2;  // / ((W-1)/2), / ((H-1)/2)
};
struct RowTap {  // 16 bytes
    int i0, i1;  // T8 rows
    float l0, l1;
};
constexpr int UP_TROWS = 8;  // low-res rows under a 48-row tile: at most 48 * (h-1)/(8h-1) + 2 < 8

template <int TW, int TH, int BW, int BH, int MINB, int WC, int HC, bool WRITE>
__global__ void __launch_bounds__(256, MINB) fbbox_up_kernel(const __grid_constant__ CUtensorMap tm_other,
                                                             const __grid_constant__ CUtensorMap tp_other, UpArgs a) {
    constexpr int DM = DM_FAST;
    constexpr int NX = TW / 32, NR = TH / 8;
    static_assert(TW % 32 == 0 && TH % 16 == 0 && 256 % TW == 0 && (256 / TW) * 2 == UP_TROWS && TH <= 256 && TH / 8 + 2 <= UP_TROWS, "tile shape");
    constexpr uint32_t BOX_BYTES = 2 * BW * BH * 4;
    extern __shared__ __align__(1024) uint8_t fbt_smem[];
    F2* T8 = reinterpret_cast<F2*>(fbt_smem + BOX_BYTES);                           // [UP_TROWS][TW]
    RowTap* rowtap = reinterpret_cast<RowTap*>(fbt_smem + BOX_BYTES + UP_TROWS * TW * 8);  // [TH]
    __shared__ uint64_t full;
    __shared__ int2 origin;
    __shared__ int nglobal_cta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {  // fold every derived constant
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw2 = const_div<DM>((float)(WC - 1) / 2.0f);
        a.dh2 = const_div<DM>((float)(HC - 1) / 2.0f);
    }
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int b = blockIdx.z;
    const float* lo = a.lo + b * a.lo_stride;
    if (threadIdx.x == 0) {
        nglobal_cta = 0;
        tc::mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 32 && b + a.pf_samples < a.B)  // the same tile of the gather source of a LATER sample -> L2
        tma_prefetch_3d(&tp_other, tx0, ty0, (b + a.pf_samples) * 2);
    if (warp == 0) {  // place the box (own flow at the lattice points straight from the low-res link) and start the copy
        const int X = tx0 + ((lane & 7) * (TW - 1)) / 7, Y = ty0 + ((lane >> 3) * (TH - 1)) / 3;
        const UpLink L{lo, a.h, a.w, a.rh, a.rw};
        const float2 v = L.value(Y, X);
        const float fnx = norm_flow_h(v.x, a.dw2), fny = norm_flow_h(v.y, a.dh2);
        const float c1x = add(norm_coord_h((float)X, a.dw2), fnx), c1y = add(norm_coord_h((float)Y, a.dh2), fny);
        const bool inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);
        const int x0 = __float2int_rd(mul(add(c1x, 1.0f), a.half_w)), y0 = __float2int_rd(mul(add(c1y, 1.0f), a.half_h));
        const int mnx = __reduce_min_sync(0xffffffffu, inb ? x0 : INT_MAX);
        const int mxx = __reduce_max_sync(0xffffffffu, inb ? x0 : INT_MIN);
        const int mny = __reduce_min_sync(0xffffffffu, inb ? y0 : INT_MAX);
        const int mxy = __reduce_max_sync(0xffffffffu, inb ? y0 : INT_MIN);
        if (lane == 0) {
            int ox, oy;
            if (mnx == INT_MAX) {
                ox = tx0 - (BW - TW) / 2;
                oy = ty0 - (BH - TH) / 2;
            } else {
                ox = mnx - (BW - (mxx + 2 - mnx)) / 2;
                oy = mny - (BH - (mxy + 2 - mny)) / 2;
            }
            ox &= ~3;  // TMA: 16-byte aligned box start
            origin = make_int2(ox, oy);
            mbar_arrive_expect_tx(&full, BOX_BYTES);
            tma_load_3d(fbt_smem, &tm_other, ox, oy, b * 2, &full);
        }
        __syncwarp();
    }
    // ---- own-flow tables (while the box is in flight) ---------------------------------------------------------------
    {
        const int r_lo = axis_tap(ty0, a.rh, a.h).i0;
        const int c = threadIdx.x & (TW - 1), rg = threadIdx.x / TW;
        const AxisTap tx = axis_tap(min(tx0 + c, W - 1), a.rw, a.w);
        const int hw = a.h * a.w;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int r = 2 * rg + u;
            const float* p = lo + min(r_lo + r, a.h - 1) * a.w;
            const float ax = mul(8.0f, __ldg(p + tx.i0)), bx = mul(8.0f, __ldg(p + tx.i1));
            const float ay = mul(8.0f, __ldg(p + hw + tx.i0)), by = mul(8.0f, __ldg(p + hw + tx.i1));
            T8[r * TW + c] = pk(fma_(tx.l0, ax, mul(tx.l1, bx)), fma_(tx.l0, ay, mul(tx.l1, by)));
        }
        if (threadIdx.x < TH) {
            const AxisTap ty = axis_tap(min(ty0 + (int)threadIdx.x, H - 1), a.rh, a.h);
            RowTap t;
            t.i0 = ty.i0 - r_lo; t.i1 = ty.i1 - r_lo; t.l0 = ty.l0; t.l1 = ty.l1;
            rowtap[threadIdx.x] = t;
        }
    }
    __syncthreads();
    const Div2 dW = make_div2(a.dw2), dH = make_div2(a.dh2);
    const F2 one2 = pk1(1.0f), hw2 = pk1(a.half_w), hh2 = pk1(a.half_h), a1_2 = pk1(a.a1), a2_2 = pk1(a.a2);
    const int X0 = tx0 + lane, Y0 = ty0 + warp;
    uint8_t* mp = a.mask + (int64_t)b * HW + Y0 * W + X0;
    float* op = WRITE ? a.own_out + (int64_t)b * 2 * HW + Y0 * W + X0 : nullptr;
    const int rstep = 8 * W;
    F2 xn2[NX];
#pragma unroll
    for (int c = 0; c < NX; c++) xn2[c] = pk1(norm_coord_h((float)(X0 + 32 * c), a.dw2));
    const uint32_t t8_s = tc::smem_u32(T8) + lane * 8, rowtap_s = tc::smem_u32(rowtap) + warp * 16;
    mbar_wait_bounded(&full, 0);
    const int2 o = origin;
    const float* sp = reinterpret_cast<const float*>(fbt_smem);
    int nglobal = 0;
    const float* g = a.other + (int64_t)b * 2 * HW;
#pragma unroll(NR / 2)
    for (int kp = 0; kp < NR; kp += 2) {
        RowTap rt[2];
#pragma unroll
        for (int p = 0; p < 2; p++)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(rt[p].i0), "=r"(rt[p].i1), "=f"(rt[p].l0), "=f"(rt[p].l1)
                         : "r"(rowtap_s + (kp + p) * 8 * 16));
        const F2 yn = sub2(dH(pk((float)(Y0 + 8 * kp), (float)(Y0 + 8 * kp + 8))), one2);
#pragma unroll
        for (int c = 0; c < NX; c++) {
            float fxs[2], fys[2];
#pragma unroll
            for (int p = 0; p < 2; p++) {  // U(Y, X) = fma(l0y, T8(i0y, X), l1y * T8(i1y, X)), both channels
                F2 t0, t1;
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t0.v) : "r"(t8_s + (rt[p].i0 * TW + 32 * c) * 8));
                asm volatile("ld.shared.b64 %0, [%1];" : "=l"(t1.v) : "r"(t8_s + (rt[p].i1 * TW + 32 * c) * 8));
                unpk(fma2(pk1(rt[p].l0), t0, mul2(pk1(rt[p].l1), t1)), fxs[p], fys[p]);
                if (WRITE) {
                    float* q = ptr_at(op, (kp + p) * rstep + 32 * c);
                    q[0] = fxs[p];
                    *ptr_at(q, HW) = fys[p];
                }
            }
            fb_pair<BW, BH>(fxs, fys, xn2[c], yn, dW, dH, one2, hw2, hh2, a1_2, a2_2, sp, o, g, W, H, HW,
                            byte_ptr_at(mp, kp * rstep + 32 * c), byte_ptr_at(mp, (kp + 1) * rstep + 32 * c), nglobal);
        }
    }
    const int wsum = __reduce_add_sync(0xffffffffu, nglobal);
    if (lane == 0 && wsum) atomicAdd(&nglobal_cta, wsum);
    __syncthreads();
    if (threadIdx.x == 0 && nglobal_cta) atomicAdd(&g_redo_pixels, (unsigned long long)nglobal_cta);
}

// ---------------------------------------------------------------------------------------------
// Dense-link chain (contrast/util.py:301-330, num > 1) with each link's gather footprint staged by
// TMA.  Same idea as fbbox_kernel, but the gather position of step i depends on steps < i, so per
// step: every thread evaluates the tap origin of its 12 pixels, the CTA reduces the exact bounding
// box (warp redux + 4 shared atomics + one __syncthreads), thread 0 centres the BW x BH box on it
// and issues the TMA copy of link i, and the taps are read from shared memory.  Positions are
// recomputed after the wait instead of being kept in registers (12 pixels x 4 values would not
// fit).  Taps outside the frame are TMA zero fill (= grid_sample's zero padding); a pixel whose
// footprint misses the box reads global memory in line.  Arithmetic = chain_kernel<false,false>.
// grid = (W / TW, H / TH, B * ndir); block = 256.
// PIXPRO_B200_CHAINBOX: 0 = gather kernels only, 1 (default) = TMA-staged dense chain where it pays
static int chainbox_mode() {
    static const int m = [] { const char* e = getenv("PIXPRO_B200_CHAINBOX"); return e ? atoi(e) : 1; }();
    return m;
}

// order-preserving float <-> int (for integer warp / shared-memory min and max of floats)
__device__ __forceinline__ int fkey(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float funkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// floor of the un-normalised sampling coordinate of pixel coordinate c (util.py:334-339 + GridSampler.cuh:22-31)
__device__ __forceinline__ int tap_origin(float c, const Div<DM_FAST>& d, float half) {
    return __float2int_rd(mul(add(norm_coord(c, d), 1.0f), half));
}

struct ChainBoxArgs {
    const float* links[2];  // per direction: link i of sample b at links + b * stride_b + i * stride_n, each [2,H,W]
    float* out[2];          // per direction, [B,2,H,W]
    int plane_n, plane_b;   // stride_n / (H*W), stride_b / (H*W)
    int n, H, W, ndir;
    float half_w, half_h;
    Div<DM_FAST> dw, dh;
};

template <int TW, int TH, int BW, int BH, int MINB>
__global__ void __launch_bounds__(256, MINB) chainbox_kernel(const __grid_constant__ CUtensorMap tm0,
                                                             const __grid_constant__ CUtensorMap tm1, ChainBoxArgs a) {
    constexpr int NX = TW / 32, NR = TH / 8;
    constexpr uint32_t BOX_BYTES = 2 * BW * BH * 4;
    extern __shared__ __align__(1024) uint8_t fbt_smem[];
    __shared__ uint64_t full;
    __shared__ int bb[4];  // min x0, max x0, min y0, max y0 of the taps that touch the frame
    __shared__ int2 origin;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = a.W, H = a.H, HW = H * W;
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const int X0 = blockIdx.x * TW + lane, Y0 = blockIdx.y * TH + warp;
    if (threadIdx.x == 0) {
        bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN;  // fkey space
        tc::mbar_init(&full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const Div2 dW = make_div2(a.dw), dH = make_div2(a.dh);
    const F2 one2 = pk1(1.0f), two2 = pk1(2.0f), hw2 = pk1(a.half_w), hh2 = pk1(a.half_h);
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    // chain state: pixel (c, k) -> (X0 + 32 c, Y0 + 8 k); pairs (k, k+1) share a packed register
    F2 cx[NX][NR / 2], cy[NX][NR / 2];
#pragma unroll
    for (int c = 0; c < NX; c++)
#pragma unroll
        for (int k = 0; k < NR / 2; k++) {
            cx[c][k] = pk1((float)(X0 + 32 * c));
            cy[c][k] = pk((float)(Y0 + 16 * k), (float)(Y0 + 16 * k + 8));
        }
    const float* sp = reinterpret_cast<const float*>(fbt_smem);
    for (int i = 0; i < a.n; i++) {
        // ---- bounding box of the tap origins ---------------------------------------------------------
        // x0 = floor(unnormalise(normalise(cx))) is a monotone function of cx (every step is), so the box of
        // the origins is that function of the box of the coordinates: 4 min/max per pixel here, the exact
        // arithmetic once per thread below.  Coordinates are clamped to where a 2x2 footprint still
        // touches the frame (pixels further out sample zeros whatever the box holds); NaNs drop out.
        float fmnx = 3.0e38f, fmxx = -3.0e38f, fmny = 3.0e38f, fmxy = -3.0e38f;
#pragma unroll
        for (int c = 0; c < NX; c++)
#pragma unroll
            for (int k = 0; k < NR / 2; k++) {
                float xs[2], ys[2];
                unpk(cx[c][k], xs[0], xs[1]);
                unpk(cy[c][k], ys[0], ys[1]);
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const float x = fminf(fmaxf(xs[p], -1.0f), wm1 + 0.5f), y = fminf(fmaxf(ys[p], -1.0f), hm1 + 0.5f);
                    fmnx = fminf(fmnx, x); fmxx = fmaxf(fmxx, x); fmny = fminf(fmny, y); fmxy = fmaxf(fmxy, y);
                }
            }
        {
            const int k0 = __reduce_min_sync(0xffffffffu, fkey(fmnx)), k1 = __reduce_max_sync(0xffffffffu, fkey(fmxx));
            const int k2 = __reduce_min_sync(0xffffffffu, fkey(fmny)), k3 = __reduce_max_sync(0xffffffffu, fkey(fmxy));
            if (lane == 0) { atomicMin(&bb[0], k0); atomicMax(&bb[1], k1); atomicMin(&bb[2], k2); atomicMax(&bb[3], k3); }
        }
        __syncthreads();  // also: every thread is done reading the previous link's box
        const float* lp = nullptr;
        int2 o;
        {
            if (threadIdx.x == 0) {
                const int bx0 = tap_origin(funkey(bb[0]), a.dw, a.half_w), bx1 = tap_origin(funkey(bb[1]), a.dw, a.half_w);
                const int by0 = tap_origin(funkey(bb[2]), a.dh, a.half_h), by1 = tap_origin(funkey(bb[3]), a.dh, a.half_h);
                int ox = bx0 - (BW - (bx1 + 2 - bx0)) / 2, oy = by0 - (BH - (by1 + 2 - by0)) / 2;
                ox &= ~3;  // TMA: 16-byte aligned box start
                origin = make_int2(ox, oy);
                bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN;  // for the next link (ordered by the barrier below)
                mbar_arrive_expect_tx(&full, BOX_BYTES);
                tma_load_3d(fbt_smem, dir ? &tm1 : &tm0, ox, oy, b * a.plane_b + i * a.plane_n, &full);
            }
            mbar_wait_bounded(&full, i & 1);
            o = origin;
            lp = (dir ? a.links[1] : a.links[0]) + ((int64_t)b * a.plane_b + (int64_t)i * a.plane_n) * HW;
        }
        // ---- gather and advance ----------------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < NX; c++)
#pragma unroll
            for (int k = 0; k < NR / 2; k++) {
                const F2 gx = sub2(dW(mul2(two2, cx[c][k])), one2), gy = sub2(dH(mul2(two2, cy[c][k])), one2);
                // products as fma(a, b, +0): an fma result cannot be contracted into the subtraction below;
                // identical to the rounded product unless it is -0, which (g + 1) * half never is here
                const F2 ix = mul2_nc(add2(gx, one2), hw2), iy = mul2_nc(add2(gy, one2), hh2);   // GridSampler.cuh:22-31
                float ixs[2], iys[2], xws[2], yns[2], t[2][8];
                unpk(ix, ixs[0], ixs[1]); unpk(iy, iys[0], iys[1]);
                bool touches[2], miss[2];
                int x0s[2], y0s[2];
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const int x0 = __float2int_rd(ixs[p]), y0 = __float2int_rd(iys[p]);  // saturates for huge, 0 for NaN
                    x0s[p] = x0; y0s[p] = y0;
                    xws[p] = __int2float_rn(x0); yns[p] = __int2float_rn(y0);
                    // the 2x2 footprint touches the frame: -1 <= x0 <= W-1 and -1 <= y0 <= H-1
                    touches[p] = (unsigned)(x0 + 1) <= (unsigned)W && (unsigned)(y0 + 1) <= (unsigned)H;
                    const unsigned dx = (unsigned)(x0 - o.x), dy = (unsigned)(y0 - o.y);
                    miss[p] = touches[p] && (dx > (unsigned)(BW - 2) || dy > (unsigned)(BH - 2));
                    const float* q = sp + min(dy, (unsigned)(BH - 2)) * BW + min(dx, (unsigned)(BW - 2));
                    t[p][0] = q[0]; t[p][1] = q[1]; t[p][2] = q[BW]; t[p][3] = q[BW + 1];
                    t[p][4] = q[BW * BH]; t[p][5] = q[BW * BH + 1]; t[p][6] = q[BW * BH + BW]; t[p][7] = q[BW * BH + BW + 1];
                }
                if (miss[0] || miss[1]) {  // rare: footprint outside the staged box, predicated taps from global memory
#pragma unroll
                    for (int p = 0; p < 2; p++)
                        if (miss[p]) {
                            const int x0 = x0s[p], y0 = y0s[p];
                            const bool x0in = x0 >= 0, x1in = x0 + 1 < W, y0in = y0 >= 0, y1in = y0 + 1 < H;
#pragma unroll
                            for (int u = 0; u < 8; u++) t[p][u] = 0.0f;
                            const float* g = lp + y0 * W + x0;
                            if (x0in && y0in) { t[p][0] = __ldg(g); t[p][4] = __ldg(g + HW); }
                            if (x1in && y0in) { t[p][1] = __ldg(g + 1); t[p][5] = __ldg(g + HW + 1); }
                            if (x0in && y1in) { t[p][2] = __ldg(g + W); t[p][6] = __ldg(g + HW + W); }
                            if (x1in && y1in) { t[p][3] = __ldg(g + W + 1); t[p][7] = __ldg(g + HW + W + 1); }
                        }
                }
                const F2 xw2 = pk(xws[0], xws[1]), yn2 = pk(yns[0], yns[1]);
                const F2 wx = sub2(ix, xw2), e = sub2(add2(xw2, one2), ix), wy = sub2(iy, yn2), s_ = sub2(add2(yn2, one2), iy);
                const F2 nw = mul2(s_, e), ne = mul2(s_, wx), sw = mul2(wy, e), se = mul2(wy, wx);
                float sxs[2], sys[2];
                unpk(combine4_2(pk(t[0][0], t[1][0]), pk(t[0][1], t[1][1]), pk(t[0][2], t[1][2]), pk(t[0][3], t[1][3]), nw, ne, sw, se), sxs[0], sxs[1]);
                unpk(combine4_2(pk(t[0][4], t[1][4]), pk(t[0][5], t[1][5]), pk(t[0][6], t[1][6]), pk(t[0][7], t[1][7]), nw, ne, sw, se), sys[0], sys[1]);
                // a footprint entirely outside the frame samples only padding: +0 (its weights may be garbage)
                const F2 sx = pk(touches[0] ? sxs[0] : 0.0f, touches[1] ? sxs[1] : 0.0f);
                const F2 sy = pk(touches[0] ? sys[0] : 0.0f, touches[1] ? sys[1] : 0.0f);
                cx[c][k] = add2(cx[c][k], sx);  // util.py:323
                cy[c][k] = add2(cy[c][k], sy);
            }
    }
    float* op = (dir ? a.out[1] : a.out[0]) + (int64_t)b * 2 * HW + Y0 * W + X0;
#pragma unroll
    for (int c = 0; c < NX; c++)
#pragma unroll
        for (int k = 0; k < NR / 2; k++) {
            float ox[2], oy[2];
            unpk(sub2(cx[c][k], pk1((float)(X0 + 32 * c))), ox[0], ox[1]);  // util.py:328
            unpk(sub2(cy[c][k], pk((float)(Y0 + 16 * k), (float)(Y0 + 16 * k + 8))), oy[0], oy[1]);
#pragma unroll
            for (int p = 0; p < 2; p++) {
                float* q = ptr_at(op, (16 * k + 8 * p) * W + 32 * c);
                q[0] = ox[p];
                *ptr_at(q, HW) = oy[p];
            }
        }
}

// ---- host side ------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// [planes, H, W] fp32 planes viewed as a 3-D tensor (W fastest); box = BW x BH x 2 planes.
static bool make_map(CUtensorMap* tm, const float* base, int64_t planes, int H, int W, int BW, int BH) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
    cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int TW, int TH, int BW, int BH, int MINB, int WC = 0, int HC = 0>
static int launch_cfg(const Args& a, int64_t B, cudaStream_t st) {
    auto kern = fbbox_kernel<TW, TH, BW, BH, MINB, WC, HC>;
    constexpr int smem = 2 * BW * BH * 4;
    static unsigned long long opted = 0;  // one bit per device
    static const bool dbg = getenv("PIXPRO_B200_FBDBG") != nullptr;
    if (a.W % TW != 0 || a.H % TH != 0) return -1;
    {
        cudaError_t e = smem_opt_in(kern, smem, opted);
        if (e != cudaSuccess) {
            if (dbg) fprintf(stderr, "fbbox: cudaFuncSetAttribute(%d) failed: %s\n", smem, cudaGetErrorString(e));
            cudaGetLastError();
            return -1;
        }
    }
    CUtensorMap tm0, tm1, tp0, tp1;
    if (!make_map(&tm0, a.flow[0], B * 2, a.H, a.W, BW, BH) || !make_map(&tm1, a.flow[1], B * 2, a.H, a.W, BW, BH) ||
        !make_map(&tp0, a.flow[0], B * 2, a.H, a.W, TW, TH) || !make_map(&tp1, a.flow[1], B * 2, a.H, a.W, TW, TH)) {
        if (dbg) fprintf(stderr, "fbbox: tensor map encode failed\n");
        return -1;
    }
    dim3 grid(a.W / TW, a.H / TH, (unsigned)(B * a.ndir));
    PP_LAUNCH(a.ndir == 1 ? "fb1" : "fb", st, (kern<<<grid, 256, smem, st>>>(tm0, tm1, tp0, tp1, a)));  // fb1: one direction
    return check_launch("fbbox_kernel");
}

static bool variant_enabled() {  // PIXPRO_B200_FBTILE=0: gather kernels only
    static const int v = [] { const char* e = getenv("PIXPRO_B200_FBTILE"); return e ? atoi(e) : 1; }();
    return v != 0;
}

// returns -1 when the TMA path is not applicable (caller falls back to the gather kernels)
static int launch(const float* f0, const float* f1, uint8_t* m0, uint8_t* m1, int ndir, int64_t B, int H, int W, float a1, float a2,
                  cudaStream_t st) {
    if (H < 2 || W < 64 || H >= 32768 || W >= 32768 || B * ndir > 65535 || !(a1 >= 0.0f)) return -1;
    if (((uintptr_t)f0 | (uintptr_t)f1) & 15) return -1;
    Args a;
    a.flow[0] = f0; a.flow[1] = f1; a.mask[0] = m0; a.mask[1] = m1;
    a.H = H; a.W = W; a.ndir = ndir;
    a.B = (int)B;
    static const int pf = [] { const char* e = getenv("PIXPRO_B200_FBPF"); return e ? atoi(e) : 1; }();
    a.pf_samples = pf > 0 ? pf : (int)B;  // >= B disables the prefetch
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.a1 = a1; a.a2 = a2;
    a.dw = make_div<DM_FAST>((float)(W - 1)); a.dh = make_div<DM_FAST>((float)(H - 1));
    a.dw2 = make_div<DM_FAST>((float)(W - 1) / 2.0f); a.dh2 = make_div<DM_FAST>((float)(H - 1) / 2.0f);
    static const int variant = [] { const char* e = getenv("PIXPRO_B200_FBTILE"); return e ? atoi(e) : 1; }();
    if (variant == 0) return -1;  // disabled: gather kernels
    // the published frame size: every derived constant folds into instruction immediates.  Box row pitch: 96 floats = 0 mod 32
    // banks measured best (360 us); pitches 88 / 92 / 100 / 104 / 108 (with the box height adjusted to the same shared memory)
    // trade same-row conflicts for cross-row ones and cost 376-412 us (profiles/r02_t_fb_pitch.txt).
    // measured and removed again (profiles/r02_u_flow_overlap.txt, r02_ab_fb_boxrows.txt): 3 CTAs / SM at 72 registers (10 % slower);
    // 64- and 60-row boxes (359 / 358 us vs 361 us, 0.7 % / 1.4 % of the pixels on the global path: free — fbbox_up_kernel uses 64 rows)
    if (W == 1280 && H == 720) return launch_cfg<64, 48, 96, 72, 4, 1280, 720>(a, B, st);
    if (H % 48 == 0) return launch_cfg<64, 48, 96, 72, 4>(a, B, st);
    if (H % 32 == 0) return launch_cfg<64, 32, 96, 56, 4>(a, B, st);
    return -1;
}

template <int TW, int TH, int BW, int BH, int MINB, int WC, int HC, bool WRITE>
static int launch_up_cfg(const UpArgs& a, cudaStream_t st) {
    auto kern = fbbox_up_kernel<TW, TH, BW, BH, MINB, WC, HC, WRITE>;
    constexpr int smem = 2 * BW * BH * 4 + UP_TROWS * TW * 8 + TH * 16;
    static unsigned long long opted = 0;  // one bit per device
    if (smem_opt_in(kern, smem, opted) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    CUtensorMap tm, tp;
    if (!make_map(&tm, a.other, (int64_t)a.B * 2, a.H, a.W, BW, BH) || !make_map(&tp, a.other, (int64_t)a.B * 2, a.H, a.W, TW, TH)) return -1;
    dim3 grid(a.W / TW, a.H / TH, (unsigned)a.B);
    PP_LAUNCH(WRITE ? "fb_up_w" : "fb_up", st, (kern<<<grid, 256, smem, st>>>(tm, tp, a)));
    return check_launch("fbbox_up_kernel");
}

// PIXPRO_B200_FBUP: n == 1 flow stages with masks — 0 = upchain1 (both composites) + fbbox (both masks); 1 = upchain1
// (forward) + fbbox_up<WRITE> (backward) + fbbox_up (forward); 2 = the same with the forward mask on fbbox_kernel.
// Unset: by batch size.  Measured (profiles/r02_ag_flow_routes.txt, one call, graph replay, B200): B = 64: 505.9 / 491.6 /
// 485.4 us for modes 0 / 1 / 2; B = 32: 264.2 / 258.1 / 260.1 us; B = 16: 141.3 / 143.4 / 147.5 us — three dependent launches
// need more waves than two to amortise their tails, so small batches (the 16-sample chunks of HostPixelStep) stay on mode 0.
static int up_mode(int64_t B) {
    static const int m = [] { const char* e = getenv("PIXPRO_B200_FBUP"); return e ? atoi(e) : -1; }();
    return m >= 0 ? m : (B >= 32 ? 2 : 0);
}
// shapes the fused kernel handles: whole 64 x 48 tiles, x8 up-sampling (the table sizes assume it), TMA-addressable planes
static bool up_applicable(const float* c0, const float* c1, int64_t B, int H, int W, int h, int w, float a1) {
    return up_mode(B) != 0 && variant_enabled() && H == 8 * h && W == 8 * w && h >= 2 && w >= 2 && W % 64 == 0 && H % 48 == 0 && H < 32768 &&
           W < 32768 && B <= 65535 && a1 >= 0.0f && ((((uintptr_t)c0 | (uintptr_t)c1) & 15) == 0) && encode_tiled_fn() != nullptr;
}
// FB mask of one direction, own composite computed from the low-res link (and written to own_out when it is not null).
// -1 = not applicable (only possible when up_applicable() was not checked, or a tensor map could not be encoded).
static int launch_up(const float* lo, int64_t lo_stride, float* own_out, const float* other, uint8_t* mask, int64_t B, int H, int W,
                     int h, int w, float a1, float a2, cudaStream_t st) {
    UpArgs a;
    a.lo = lo; a.lo_stride = lo_stride; a.own_out = own_out; a.other = other; a.mask = mask;
    a.H = H; a.W = W; a.h = h; a.w = w; a.B = (int)B;
    static const int pf = [] { const char* e = getenv("PIXPRO_B200_FBPF"); return e ? atoi(e) : 1; }();
    a.pf_samples = pf > 0 ? pf : (int)B;
    a.rh = up_scale(h, H); a.rw = up_scale(w, W);
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.a1 = a1; a.a2 = a2;
    a.dw2 = make_div<DM_FAST>((float)(W - 1) / 2.0f); a.dh2 = make_div<DM_FAST>((float)(H - 1) / 2.0f);
    if (W == 1280 && H == 720)
        return own_out ? launch_up_cfg<64, 48, 96, 64, 4, 1280, 720, true>(a, st) : launch_up_cfg<64, 48, 96, 64, 4, 1280, 720, false>(a, st);
    return own_out ? launch_up_cfg<64, 48, 96, 64, 4, 0, 0, true>(a, st) : launch_up_cfg<64, 48, 96, 64, 4, 0, 0, false>(a, st);
}

template <int TW, int TH, int BW, int BH, int MINB>
static int launch_chain_cfg(const ChainBoxArgs& a, int64_t B, const float* base0, const float* base1, int64_t planes0,
                            int64_t planes1, cudaStream_t st) {
    auto kern = chainbox_kernel<TW, TH, BW, BH, MINB>;
    constexpr int smem = 2 * BW * BH * 4;
    static unsigned long long opted = 0;  // one bit per device
    if (a.W % TW != 0 || a.H % TH != 0) return -1;
    if (smem_opt_in(kern, smem, opted) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    CUtensorMap tm0, tm1;
    memset(&tm0, 0, sizeof(tm0));
    if (!make_map(&tm0, base0, planes0, a.H, a.W, BW, BH)) return -1;
    if (a.ndir == 2) {
        if (!make_map(&tm1, base1, planes1, a.H, a.W, BW, BH)) return -1;
    } else {
        tm1 = tm0;
    }
    dim3 grid(a.W / TW, a.H / TH, (unsigned)(B * a.ndir));
    PP_LAUNCH("chain_dense", st, (kern<<<grid, 256, smem, st>>>(tm0, tm1, a)));
    return check_launch("chainbox_kernel");
}

// Dense-link chain through the TMA-staged kernel; -1 = not applicable (caller uses the gather kernels).
static int launch_chain_box(const float* l0, const float* l1, float* o0, float* o1, int ndir, int n, int64_t B, int H, int W,
                            int64_t stride_n, int64_t stride_b, cudaStream_t st) {
    if (chainbox_mode() < 1 || n < 2 || H < 2 || W < 64 || H >= 32768 || W >= 32768 || B * ndir > 65535) return -1;
    // one CTA per 64x48 tile and link-by-link TMA round trips: pays off from ~3 waves of CTAs (measured: 1.2-1.3x
    // faster than the gather kernel at 8-32 planes of 720x1280, slower for the 2-plane chunks of pp_flow_stage)
    if ((int64_t)(W / 64) * (H / 48) * B * ndir < 1200) return -1;
    const int64_t HW = (int64_t)H * W;
    if (stride_n % HW != 0 || stride_b % HW != 0) return -1;
    if (((uintptr_t)l0 | (uintptr_t)(ndir == 2 ? l1 : l0)) & 15) return -1;
    ChainBoxArgs a;
    a.links[0] = l0; a.links[1] = l1; a.out[0] = o0; a.out[1] = o1;
    a.plane_n = (int)(stride_n / HW); a.plane_b = (int)(stride_b / HW);
    a.n = n; a.H = H; a.W = W; a.ndir = ndir;
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.dw = make_div<DM_FAST>((float)(W - 1)); a.dh = make_div<DM_FAST>((float)(H - 1));
    // planes addressable from each base: the last link of the last sample ends at this plane
    const int64_t planes = (B - 1) * a.plane_b + (int64_t)(n - 1) * a.plane_n + 2;
    return launch_chain_cfg<64, 48, 96, 72, 4>(a, B, l0, l1, planes, planes, st);
}

}  // namespace fbt
}  // namespace pp
