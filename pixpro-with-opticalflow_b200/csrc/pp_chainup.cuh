// pp_chainup.cuh — x8 up-sampling fused with the chain of n > 1 links (contrast/flow/utils/utils.py:87-89 +
// contrast/util.py:301-330, the `flow_up` path of apply_optical_flow for n_frames > 2), ONE launch for both
// directions of the whole batch and NO full-resolution scratch: HBM traffic = the low-res links in, the
// composites out (n*230 400 + 14 745 600 B per sample).  Included by pp_flow.cu.
//
// The reference materialises every up-sampled link ([n*B,2,720,1280], 4.7 GB per direction at B=128) and
// grid_samples it n times.  Here one CTA owns a TW x TH tile of output pixels and keeps their running
// coordinates in registers, one (x, y) pair per pixel in a packed fp32 register pair.  Per link:
//   A. the tile's tap footprint is bounded exactly from the running coordinates (warp redux + 4 shared atomics);
//   B. the low-res patch under the footprint (<= 24 x 24 values per channel, pre-scaled by the up-sampling's
//      factor 8) comes in by ONE round trip to L2, and the axis taps (ATen's area_pixel_compute_source_index,
//      UpSample.cuh) of the footprint's columns and rows are computed once, by one thread each, while it is in flight;
//   C. the low-res rows under the footprint are interpolated HORIZONTALLY at the footprint's columns into shared
//      memory: T8(r, X) = fma(l0x, 8 L[r][i0x], l1x * 8 L[r][i1x]), both channels as one packed pair;
//   D. every pixel evaluates the up-sampled link at its four integer taps from T8,
//      U(Y, X) = fma(l0y, T8(i0y, X), l1y * T8(i1y, X)) — bit-identical to ATen's
//      8 * fma(l0y, fma(l0x,a,l1x*b), l1y * fma(l0x,c,l1x*d)) because scaling by 8 is exact and commutes with
//      every rounding (no intermediate is subnormal for flows above 2^-120 px) — combines them with grid_sample's
//      own arithmetic and advances.  The full-resolution field never exists, not even in shared memory: chained
//      flows stretch a tile's footprint (measured on the bench's fields: 76 x 65 px median, 141 x 148 px at most for
//      a 64 x 48 tile at the fifth link), and a footprint-sized box of it would cost more to fill than the taps
//      cost to evaluate (the first version of this kernel did that: 183 instructions per pixel-link, 45 % of the
//      tiles needing a second box).  T8 holds up to 160 columns x 24 low-res rows = a 160 x 168 px footprint.
// Rows / columns outside the frame hold 0 = grid_sample's zero padding.  A pixel whose footprint is not inside the
// staged region (none on the bench's fields) evaluates its taps from the low-res link in global memory
// (UpLink::value — same values).  Arithmetic = chain_kernel<true,false,DM_FAST>.
#pragma once
#include "pp_common.cuh"

namespace pp {
namespace cup {

struct Args {
    const float* links[2];  // per direction: link i of sample b at links + b * stride_b + i * stride_n, each [2,h,w]
    float* out[2];          // per direction, [B,2,H,W]
    int64_t stride_n, stride_b;
    int n, h, w, H, W, ndir;
    float rh, rw, half_w, half_h;
    Div<DM_FAST> dw, dh;
};

__device__ unsigned long long g_slow_pixels;  // pixel-links that took the direct (slow) path: diagnostics

struct AxTap {  // 16 bytes
    int i0, i1;  // rows: T8 row indices (the zero row for rows outside the frame); columns: patch columns, i0 < 0 = outside
    float l0, l1;
};

constexpr int TC = 160;            // staged columns (full-res)
constexpr int TRW = 24;            // staged low-res rows; row TRW of T8 is all zeros (rows outside the frame point there)
constexpr int FH_MAX = (TRW - 3) * 8;  // staged full-res rows: they touch at most FH_MAX / 8 + 3 low-res rows
constexpr int PW = TC / 8 + 4;     // staged low-res columns

constexpr int smem_bytes() {
    //     T8 [TRW + 1][TC] pairs   rowtap [FH_MAX + 1]   coltap [TC]   patch [2][TRW][PW]
    return (TRW + 1) * TC * 8 + (FH_MAX + 1) * 16 + TC * 16 + 2 * TRW * PW * 4;
}

// Shared-memory loads through explicit 32-bit shared addresses: with generic pointers the compiler re-derives the shared
// window base (S2UR + UMOV + ULEA) next to every use in the hot loop (ncu: 9 % of the issue slots were uniform-datapath
// glue).  `volatile`: never reordered across the barriers (which are volatile too), never deleted.
__device__ __forceinline__ F2 lds64(uint32_t addr) {
    F2 r;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v) : "r"(addr));
    return r;
}
__device__ __forceinline__ AxTap lds_tap(uint32_t addr) {
    AxTap t;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.i0), "=r"(t.i1), "=f"(t.l0), "=f"(t.l1) : "r"(addr));
    return t;
}

// The four taps of the up-sampled link at (x0, y0) .. (x0+1, y0+1) straight from the low-res field in global memory
// (last-resort path of a pixel whose footprint the staged region does not cover): out of line, off the hot loop.
__device__ __noinline__ void direct_taps(const float* lo, int h, int w, float rh, float rw, int W, int H, int x0, int y0, float t[8]) {
    const bool x0in = x0 >= 0, x1in = x0 + 1 < W, y0in = y0 >= 0, y1in = y0 + 1 < H;
#pragma unroll
    for (int u = 0; u < 8; u++) t[u] = 0.0f;
    const UpLink L{lo, h, w, rh, rw};
    if (x0in && y0in) { const float2 v = L.value(y0, x0); t[0] = v.x; t[4] = v.y; }
    if (x1in && y0in) { const float2 v = L.value(y0, x0 + 1); t[1] = v.x; t[5] = v.y; }
    if (x0in && y1in) { const float2 v = L.value(y0 + 1, x0); t[2] = v.x; t[6] = v.y; }
    if (x1in && y1in) { const float2 v = L.value(y0 + 1, x0 + 1); t[3] = v.x; t[7] = v.y; }
}

// grid = (ceil(W / TW), ceil(H / TH), B * ndir); block = 256; dynamic smem = smem_bytes().
// WC/HC > 0: the frame size is a compile-time constant (the 1280x720 frames of the published BDD100K runs): every
// derived constant folds into instruction immediates.
template <int TW, int TH, int MINB, int WC, int HC>
__global__ void __launch_bounds__(256, MINB) chainup_kernel(Args a) {
    constexpr int NX = TW / 32, NR = TH / 8, NPX = NX * NR;  // NPX pixels per thread: (X0 + 32 c, Y0 + 8 k)
    static_assert(TW % 32 == 0 && TH % 8 == 0 && TW + 2 <= TC && TH + 2 <= FH_MAX, "tile shape");
    extern __shared__ __align__(16) uint8_t cup_smem[];
    F2* T8 = reinterpret_cast<F2*>(cup_smem);                                           // [TRW + 1][TC] (x, y) pairs
    AxTap* rowtap = reinterpret_cast<AxTap*>(cup_smem + (TRW + 1) * TC * 8);              // [FH_MAX + 1]
    AxTap* coltap = rowtap + (FH_MAX + 1);                                              // [TC]
    float* patch = reinterpret_cast<float*>(coltap + TC);                               // [2][TRW][PW], already x 8
    __shared__ int bb[4];  // fkey space: min x, max x, min y, max y of the running coordinates
    __shared__ int nslow_cta;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = WC ? WC : a.W, H = HC ? HC : a.H, HW = H * W;
    if (WC) {
        a.half_w = (float)(WC - 1) / 2.0f; a.half_h = (float)(HC - 1) / 2.0f;
        a.dw = const_div<DM_FAST>((float)(WC - 1));
        a.dh = const_div<DM_FAST>((float)(HC - 1));
    }
    const int dir = a.ndir == 2 ? (blockIdx.z & 1) : 0;
    const int b = a.ndir == 2 ? (blockIdx.z >> 1) : blockIdx.z;
    const int X0 = blockIdx.x * TW + lane, Y0 = blockIdx.y * TH + warp;
    const float* lbase = (dir ? a.links[1] : a.links[0]) + b * a.stride_b;
    if (threadIdx.x == 0) {
        bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN;
        nslow_cta = 0;
    }
    for (int c = threadIdx.x; c < TC; c += 256) T8[TRW * TC + c].v = 0ull;  // the zero row
    // 2*c/s == c/(s/2) exactly (see norm_coord_h): one certified division, no doubling.  Halves: (x, y).
    const F2 INV = pk(a.dw.inv * 2.0f, a.dh.inv * 2.0f), INV_LO = pk(a.dw.inv_lo * 2.0f, a.dh.inv_lo * 2.0f);
    const F2 HALF = pk(a.half_w, a.half_h), one2 = pk1(1.0f);
    const float wm1 = (float)(W - 1), hm1 = (float)(H - 1);
    F2 cxy[NPX];  // running coordinate (x, y) of pixel j = c * NR + k
#pragma unroll
    for (int c = 0; c < NX; c++)
#pragma unroll
        for (int k = 0; k < NR; k++) cxy[c * NR + k] = pk((float)(X0 + 32 * c), (float)(Y0 + 8 * k));
    int nslow = 0;
    __syncthreads();
    for (int i = 0; i < a.n; i++) {
        const float* lo = lbase + i * a.stride_n;
        // ---- A. exact bound of the running coordinates ------------------------------------------------------
        {
            float fmnx = 3.0e38f, fmxx = -3.0e38f, fmny = 3.0e38f, fmxy = -3.0e38f;
#pragma unroll
            for (int j = 0; j < NPX; j++) {  // NaNs drop out of fminf / fmaxf
                float x, y;
                unpk(cxy[j], x, y);
                fmnx = fminf(fmnx, x); fmxx = fmaxf(fmxx, x); fmny = fminf(fmny, y); fmxy = fmaxf(fmxy, y);
            }
            // clamped to where a 2x2 footprint still touches the frame (pixels further out sample zeros whatever is staged)
            fmnx = fminf(fmaxf(fmnx, -1.0f), wm1 + 0.5f); fmxx = fminf(fmaxf(fmxx, -1.0f), wm1 + 0.5f);
            fmny = fminf(fmaxf(fmny, -1.0f), hm1 + 0.5f); fmxy = fminf(fmaxf(fmxy, -1.0f), hm1 + 0.5f);
            const int k0 = __reduce_min_sync(0xffffffffu, fbt::fkey(fmnx)), k1 = __reduce_max_sync(0xffffffffu, fbt::fkey(fmxx));
            const int k2 = __reduce_min_sync(0xffffffffu, fbt::fkey(fmny)), k3 = __reduce_max_sync(0xffffffffu, fbt::fkey(fmxy));
            if (lane == 0) { atomicMin(&bb[0], k0); atomicMax(&bb[1], k1); atomicMin(&bb[2], k2); atomicMax(&bb[3], k3); }
        }
        __syncthreads();  // (1) bb complete; every thread is done reading the previous link's tables
        // x0 = floor(unnormalise(normalise(x))) is monotone in x (every step is), so the bound of the tap origins is that
        // function of the bound of the coordinates.  Staged region = [ox, ox + fw) x [oy, oy + fh).
        const int bx0 = fbt::tap_origin(fbt::funkey(bb[0]), a.dw, a.half_w), bx1 = fbt::tap_origin(fbt::funkey(bb[1]), a.dw, a.half_w);
        const int by0 = fbt::tap_origin(fbt::funkey(bb[2]), a.dh, a.half_h), by1 = fbt::tap_origin(fbt::funkey(bb[3]), a.dh, a.half_h);
        int ox = bx0, oy = by0;
        int fw = bx1 + 2 - ox, fh = by1 + 2 - oy;
        if (fw > TC) { ox += (fw - TC) / 2; fw = TC; }          // larger than the staging area: keep the middle
        if (fh > FH_MAX) { oy += (fh - FH_MAX) / 2; fh = FH_MAX; }
        // low-res rows / columns of the part of the region that lies in the frame
        const int ya = max(oy, 0), yb = min(oy + fh - 1, H - 1), xa = max(ox, 0), xb = min(ox + fw - 1, W - 1);
        const bool any_in = ya <= yb && xa <= xb;
        const int r_lo = any_in ? axis_tap(ya, a.rh, a.h).i0 : 0, c_lo = any_in ? axis_tap(xa, a.rw, a.w).i0 : 0;
        const int nrow = any_in ? axis_tap(yb, a.rh, a.h).i1 - r_lo + 1 : 0;
        const int ncol = any_in ? axis_tap(xb, a.rw, a.w).i1 - c_lo + 1 : 0;
        // ---- B. low-res patch (one round trip to L2), axis taps -------------------------------------------------
        {
            const int hw = a.h * a.w;
            constexpr int NL = (2 * TRW * PW + 255) / 256;
            float pv[NL];
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = threadIdx.x + 256 * u;
                const int cc = e % PW, r = (e / PW) % TRW, ch = e / (PW * TRW);
                pv[u] = (e < 2 * TRW * PW && r < nrow && cc < ncol) ? __ldg(lo + ch * hw + (r_lo + r) * a.w + c_lo + cc) : 0.0f;
            }
            if (threadIdx.x < TC) {
                const int X = ox + threadIdx.x;
                AxTap t;
                t.i0 = -1; t.i1 = 0; t.l0 = 0.0f; t.l1 = 0.0f;
                if (threadIdx.x < fw && X >= 0 && X < W) {
                    const AxisTap tx = axis_tap(X, a.rw, a.w);
                    t.i0 = tx.i0 - c_lo; t.i1 = tx.i1 - c_lo; t.l0 = tx.l0; t.l1 = tx.l1;
                }
                coltap[threadIdx.x] = t;
            }
            for (int y = threadIdx.x; y <= fh; y += 256) {  // row fh: the south tap row of the last hit row + 1 never reads it, kept valid
                const int Y = oy + y;
                AxTap t;
                t.i0 = TRW; t.i1 = TRW; t.l0 = 0.0f; t.l1 = 0.0f;  // outside the frame: the zero row
                if (y < fh && Y >= 0 && Y < H) {
                    const AxisTap ty = axis_tap(Y, a.rh, a.h);
                    t.i0 = ty.i0 - r_lo; t.i1 = ty.i1 - r_lo; t.l0 = ty.l0; t.l1 = ty.l1;
                }
                rowtap[y] = t;
            }
#pragma unroll
            for (int u = 0; u < NL; u++) {
                const int e = threadIdx.x + 256 * u;
                if (e < 2 * TRW * PW) patch[e] = mul(8.0f, pv[u]);
            }
        }
        __syncthreads();  // (2) patch and taps staged; every thread has read bb
        if (threadIdx.x == 0) { bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN; }  // for the next link
        // ---- C. horizontal pass: T8[r][c] = fma(l0x, 8 L[r][i0x], l1x * 8 L[r][i1x]), both channels -----------------
        for (int c = lane; c < fw; c += 32) {
            const AxTap t = coltap[c];
            const F2 l0 = pk1(t.l0), l1 = pk1(t.l1);
            for (int r = warp; r < nrow; r += 8) {
                F2 v;
                v.v = 0ull;  // columns outside the frame hold zeros
                if (t.i0 >= 0) {
                    const float* row = patch + r * PW;
                    v = fma2(l0, pk(row[t.i0], row[TRW * PW + t.i0]), mul2(l1, pk(row[t.i1], row[TRW * PW + t.i1])));
                }
                T8[r * TC + c] = v;
            }
        }
        __syncthreads();  // (3)
        // ---- D. evaluate the taps, combine, advance ---------------------------------------------------------------
        const unsigned xlim = (unsigned)(fw - 2), ylim = (unsigned)(fh - 2);
        const uint32_t t8_s = (uint32_t)__cvta_generic_to_shared(T8), rowtap_s = (uint32_t)__cvta_generic_to_shared(rowtap);
#pragma unroll
        for (int j = 0; j < NPX; j++) {
            const F2 c = cxy[j];
            const F2 g = sub2(fma2(c, INV, mul2(c, INV_LO)), one2);                               // util.py:334-339
            // product as fma(a, b, +0): an fma result cannot be contracted into the subtraction below;
            // identical to the rounded product unless it is -0, which (g + 1) * half never is where it matters
            const F2 ii = mul2_nc(add2(g, one2), HALF);                                            // GridSampler.cuh:22-31
            float ix, iy;
            unpk(ii, ix, iy);
            const int x0 = __float2int_rd(ix), y0 = __float2int_rd(iy);  // saturates for huge, 0 for NaN
            const F2 fl = pk(__int2float_rn(x0), __int2float_rn(y0));
            // the 2x2 footprint touches the frame: -1 <= x0 <= W-1 and -1 <= y0 <= H-1
            const bool touches = (unsigned)(x0 + 1) <= (unsigned)W && (unsigned)(y0 + 1) <= (unsigned)H;
            const unsigned dx = (unsigned)(x0 - ox), dy = (unsigned)(y0 - oy);
            const bool hit = dx <= xlim && dy <= ylim;
            const unsigned dxc = min(dx, xlim), dyc = min(dy, ylim);  // clamped: always valid addresses
            const AxTap r0 = lds_tap(rowtap_s + dyc * 16), r1 = lds_tap(rowtap_s + dyc * 16 + 16);
            // T8 rows of the north taps (y0) and of the south taps (y0 + 1).  Seven times out of eight the two full-res rows
            // lie between the same two low-res rows: the south taps then reuse the loaded values (the kernel is bound by
            // shared-memory wavefronts: 8 eight-byte loads per pixel otherwise)
            const uint32_t colw = t8_s + dxc * 8;
            const F2 a0w = lds64(colw + r0.i0 * (TC * 8)), a0e = lds64(colw + r0.i0 * (TC * 8) + 8);
            const F2 a1w = lds64(colw + r0.i1 * (TC * 8)), a1e = lds64(colw + r0.i1 * (TC * 8) + 8);
            F2 b0w = a0w, b0e = a0e, b1w = a1w, b1e = a1e;
            if (r1.i0 != r0.i0 || r1.i1 != r0.i1) {
                b0w = lds64(colw + r1.i0 * (TC * 8)); b0e = lds64(colw + r1.i0 * (TC * 8) + 8);
                b1w = lds64(colw + r1.i1 * (TC * 8)); b1e = lds64(colw + r1.i1 * (TC * 8) + 8);
            }
            F2 unw = fma2(pk1(r0.l0), a0w, mul2(pk1(r0.l1), a1w));
            F2 une = fma2(pk1(r0.l0), a0e, mul2(pk1(r0.l1), a1e));
            F2 usw = fma2(pk1(r1.l0), b0w, mul2(pk1(r1.l1), b1w));
            F2 use = fma2(pk1(r1.l0), b0e, mul2(pk1(r1.l1), b1e));
            if (touches && !hit) {  // the staged region does not cover it: straight from the low-res link (rare)
                float d[8];
                direct_taps(lo, a.h, a.w, a.rh, a.rw, W, H, x0, y0, d);
                unw = pk(d[0], d[4]); une = pk(d[1], d[5]); usw = pk(d[2], d[6]); use = pk(d[3], d[7]);
                nslow++;
            }
            // east / south weights as 1 - w: for i >= 0, w = i - floor(i) is exact, so (floor + 1) - i and 1 - w round the
            // same real number; for -1 <= i < 0 the west / north taps they multiply are padding zeros; a pixel outside that
            // range takes nothing (adding the +0 the reference samples there leaves a coordinate, never -0, unchanged)
            const F2 wgt = sub2(ii, fl), est = sub2(one2, wgt);
            float wx, wy, ex, sy;
            unpk(wgt, wx, wy);
            unpk(est, ex, sy);
            const float nw = mul(sy, ex), ne = mul(sy, wx), sw = mul(wy, ex), se = mul(wy, wx);
            const F2 smp = fma2(use, pk1(se), fma2(usw, pk1(sw), fma2(une, pk1(ne), mul2(unw, pk1(nw)))));  // combine4, both channels
            const F2 nxt = add2(c, smp);  // util.py:323
            float cx_, cy_, nx_, ny_;
            unpk(c, cx_, cy_);
            unpk(nxt, nx_, ny_);
            cxy[j] = pk(touches ? nx_ : cx_, touches ? ny_ : cy_);
        }
    }
    float* op = (dir ? a.out[1] : a.out[0]) + (int64_t)b * 2 * HW + Y0 * W + X0;
#pragma unroll
    for (int c = 0; c < NX; c++)
#pragma unroll
        for (int k = 0; k < NR; k++) {
            float fx, fy;
            unpk(sub2(cxy[c * NR + k], pk((float)(X0 + 32 * c), (float)(Y0 + 8 * k))), fx, fy);  // util.py:328
            if (X0 + 32 * c < W && Y0 + 8 * k < H) {  // ragged frames: tiles may overhang
                if (WC) {
                    op[8 * k * WC + 32 * c] = fx;
                    op[WC * HC + 8 * k * WC + 32 * c] = fy;
                } else {
                    float* dst = ptr_at(op, 8 * k * W + 32 * c);
                    dst[0] = fx;
                    *ptr_at(dst, HW) = fy;
                }
            }
        }
    const int wsum = __reduce_add_sync(0xffffffffu, nslow);
    if (lane == 0 && wsum) atomicAdd(&nslow_cta, wsum);
    __syncthreads();
    if (threadIdx.x == 0 && nslow_cta) atomicAdd(&g_slow_pixels, (unsigned long long)nslow_cta);
}

static bool applicable(int n, int64_t B, int ndir, int h, int w) {
    const int H = 8 * h, W = 8 * w;
    return n >= 2 && h >= 2 && w >= 2 && H < 32768 && W < 32768 && B * ndir <= 65535;
}

template <int TW, int TH, int MINB, int WC = 0, int HC = 0>
static int launch_cfg(const Args& a, int64_t B, cudaStream_t st) {
    auto kern = chainup_kernel<TW, TH, MINB, WC, HC>;
    constexpr int smem = smem_bytes();
    static unsigned long long opted = 0;  // one bit per device
    if (smem_opt_in(kern, smem, opted) != cudaSuccess) {
        set_error("chainup_kernel: cudaFuncSetAttribute(%d B of shared memory) failed: %s", smem, cudaGetErrorString(cudaGetLastError()));
        return PP_ERR_CUDA;
    }
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, (unsigned)(B * a.ndir));
    PP_LAUNCH("chain_up", st, (kern<<<grid, 256, smem, st>>>(a)));
    return check_launch("chainup_kernel");
}

// Chain of n > 1 LOW-RES links with the x8 up-sampling fused in; -1 = not applicable.
static int launch(const float* l0, const float* l1, float* o0, float* o1, int ndir, int n, int64_t B, int h, int w,
                  int64_t stride_n, int64_t stride_b, cudaStream_t st) {
    if (!applicable(n, B, ndir, h, w)) return -1;
    const int H = 8 * h, W = 8 * w;
    Args a;
    a.links[0] = l0; a.links[1] = l1; a.out[0] = o0; a.out[1] = o1;
    a.stride_n = stride_n; a.stride_b = stride_b;
    a.n = n; a.h = h; a.w = w; a.H = H; a.W = W; a.ndir = ndir;
    a.rh = up_scale(h, H); a.rw = up_scale(w, W);
    a.half_w = (float)(W - 1) / 2.0f; a.half_h = (float)(H - 1) / 2.0f;
    a.dw = make_div<DM_FAST>((float)(W - 1)); a.dh = make_div<DM_FAST>((float)(H - 1));
    static const int variant = [] { const char* e = getenv("PIXPRO_B200_CHAINUP"); return e ? atoi(e) : 1; }();
    if (W == 1280 && H == 720) {  // the published frame size
        if (variant == 2) return launch_cfg<64, 48, 3, 1280, 720>(a, B, st);
        if (variant == 3) return launch_cfg<64, 32, 4, 1280, 720>(a, B, st);
        if (variant == 4) return launch_cfg<64, 24, 4, 1280, 720>(a, B, st);
        return launch_cfg<64, 48, 4, 1280, 720>(a, B, st);
    }
    return launch_cfg<64, 48, 4>(a, B, st);
}

}  // namespace cup
}  // namespace pp
