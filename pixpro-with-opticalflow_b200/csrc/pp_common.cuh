// pp_common.cuh — shared device helpers for the sm_100a pixel-pretext kernels.
//
// Arithmetic that decides a boolean downstream (coordinates, masks) is written with the
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, ...) so that ptxas can never contract
// two separately rounded reference ops into one FMA; FMAs appear only where the reference's
// own ATen kernels fuse (grid_sample tap accumulation, bilinear up-sampling).  See
// oracle/pixpro_oracle.c for the op-by-op restatement these helpers mirror.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pixpro_b200.h"

namespace pp {

// ---- error plumbing --------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);
void count_launch(int n = 1);

// Brackets one kernel launch with a cudaEvent pair when pp_profile_enable(1) is active
// (no-op otherwise).  Usage:  { ProfScope ps("name", st); kernel<<<...,st>>>(...); }
class ProfScope {
  public:
    ProfScope(const char* name, cudaStream_t st);
    ~ProfScope();

  private:
    const char* name_;
    cudaStream_t st_;
    cudaEvent_t e1_;
};

// One kernel launch: optional event bracket + launch accounting.
#define PP_LAUNCH(name, st, ...)               \
    do {                                       \
        ::pp::ProfScope _pp_ps(name, st);      \
        __VA_ARGS__;                           \
        ::pp::count_launch();                  \
    } while (0)

// One-time opt-in of a kernel to more than 48 KB of dynamic shared memory, PER DEVICE: the attribute belongs to the
// (function, device) pair, so a process that drives several GPUs must set it on each.  `done` is one bit mask per
// call site (i.e. per kernel instantiation), bit = device ordinal; racing first calls both set the attribute (idempotent).
template <class K>
static inline cudaError_t smem_opt_in(K kern, int bytes, unsigned long long& done) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (__atomic_load_n(&done, __ATOMIC_ACQUIRE) & bit) return cudaSuccess;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) __atomic_fetch_or(&done, bit, __ATOMIC_RELEASE);
    return e;
}

#define PP_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            ::pp::set_error(__VA_ARGS__);     \
            return PP_ERR_INVALID;            \
        }                                     \
    } while (0)

// base + off (in floats) as ONE IMAD.WIDE: nvcc otherwise widens index arithmetic on uniform
// bases into 4-instruction 64-bit add/shift sequences per load (profiles/r01_a_*).
__device__ __forceinline__ const float* ptr_at(const float* base, int off) {
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"((unsigned long long)base));
    return reinterpret_cast<const float*>(r);
}
__device__ __forceinline__ float* ptr_at(float* base, int off) {
    unsigned long long r;
    asm("mad.wide.s32 %0, %1, 4, %2;" : "=l"(r) : "r"(off), "l"((unsigned long long)base));
    return reinterpret_cast<float*>(r);
}

// Predicated read-only load with an immediate element offset, as `asm volatile`: the compiler
// keeps volatile asm statements in program order, which is what batches the gathers of several
// pixels ahead of their uses (it otherwise re-serialises them per pixel to save registers).
template <int OFF>
__device__ __forceinline__ float ldg_pred(const float* p, bool pred) {
    float v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t"
        "@q ld.global.nc.f32 %0, [%1+%3];\n\t}"
        : "=f"(v)
        : "l"(p), "r"((int)pred), "n"(OFF * 4));
    return v;
}

// ---- exactly-rounded scalar ops ----------------------------------------------------------
__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// tensor / python_scalar of the reference.  IEEE mode = correctly rounded quotient (torch
// CPU); RCP mode = x * fl32(1/s) (torch CUDA's true-divide with a CPU scalar).
struct ScalarDiv {
    float s;    // divisor
    float inv;  // fl32(1/s)
    int rcp;    // div_mode
    __device__ __forceinline__ float operator()(float x) const {
        return rcp ? __fmul_rn(x, inv) : __fdiv_rn(x, s);
    }
};
static inline ScalarDiv make_div(float s, int div_mode) {
    ScalarDiv d;
    d.s = s;
    d.inv = 1.0f / s;
    d.rcp = div_mode;
    return d;
}

// Compile-time division flavours for the flow kernels (where divisions dominate the issue
// slots).  Exact constant division in TWO instructions: with inv = fl32(1/s) and
// inv_lo = fl32(1/s - inv) (so inv + inv_lo = 1/s to ~2^-48),
//     q = fma(x, inv, RN(x * inv_lo))
// is within 2^-47 (relative) of x/s before its single final rounding, hence equals RN(x/s) unless
// x/s lies that close to a rounding boundary.  For the divisors used here (s = d or d/2 with d a
// small integer) the quotient of a 24-bit x can be no closer than 2^-25/d to a boundary, and the
// host does not rely on that argument: div_certified() checks the identity for all 2^23 mantissas
// of one binade against true division.  Every step scales exactly by powers of two as long as no
// intermediate is subnormal, so one binade covers every x in [2^-60, 2^100] ∪ {+0} (the host also
// requires inv_lo == 0 or |inv_lo| >= 2^-66, which keeps RN(x*inv_lo) normal on that range).
// Outside that range (non-finite, tiny, -0) the unguarded form may differ from IEEE division, so
// DM_FAST is used only where such inputs provably cannot reach an output bit;
// DM_FASTG adds the range guard and is exact everywhere.
enum { DM_IEEE = 0, DM_RCP = 1, DM_FAST = 2, DM_FASTG = 3 };

bool div_certified(float s);  // host, cached (pp_api.cu)

// fl32(1/s - fl32(1/s)); usable in constant expressions (device code folds it for literal s).
__host__ __device__ constexpr float recip_lo(float s) { return (float)(1.0 / (double)s - (double)(1.0f / s)); }

template <int DM>
struct Div {
    float s;       // divisor
    float inv;     // fl32(1/s)
    float inv_lo;  // fl32(1/s - inv)
    __device__ __forceinline__ float operator()(float x) const {
        if (DM == DM_RCP) return __fmul_rn(x, inv);
        if (DM == DM_IEEE) return __fdiv_rn(x, s);
        float q = __fmaf_rn(x, inv, __fmul_rn(x, inv_lo));
        if (DM == DM_FASTG) {
            float ax = fabsf(x);
            bool plain = (ax >= 8.673617379884035e-19f && ax <= 1.2676506002282294e30f) || (__float_as_uint(x) == 0u);
            if (!plain) q = __fdiv_rn(x, s);
        }
        return q;
    }
};
// Div over a compile-time divisor (every field folds to an instruction immediate).
template <int DM>
__device__ __forceinline__ Div<DM> const_div(float s) {
    Div<DM> d;
    d.s = s;
    d.inv = 1.0f / s;
    d.inv_lo = recip_lo(s);
    return d;
}
template <int DM>
static inline Div<DM> make_div(float s) {
    Div<DM> d;
    d.s = s;
    d.inv = 1.0f / s;
    d.inv_lo = recip_lo(s);
    return d;
}

// ---- packed fp32 pairs (sm_100 FFMA2 / FMUL2 / FADD2) ---------------------------------------
// One issue slot performs the same individually rounded operation on two floats.  The flow kernels
// are issue-bound on exactly-rounded scalar arithmetic, so they process two pixels per packed op.
// CAUTION (ptxas 12.9): a packed mul whose result feeds a packed add/sub IS contracted into one
// FFMA2 despite the .rn modifiers (verified in SASS; the scalar forms are never contracted).
// Wherever the reference rounds a product before adding it, the add must therefore be done with
// the scalar add()/sub() on the unpacked halves.  Other shapes (product as multiplicand or as
// fma addend, sum feeding a product) have no fused form and are safe.
struct F2 {
    unsigned long long v;
};
__device__ __forceinline__ F2 pk(float lo, float hi) {
    F2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ F2 pk1(float x) { return pk(x, x); }
__device__ __forceinline__ void unpk(F2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ F2 mul2(F2 a, F2 b) {
    F2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 add2(F2 a, F2 b) {
    F2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 sub2(F2 a, F2 b) {
    F2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}
// a*b as an FFMA2 with a +0 addend: bit-identical to the rounded product whenever the product is
// not -0 (squares, products of non-negative factors) — and, being an fma result, it can feed a
// packed add/sub without being contracted into it (see the caveat above).
__device__ __forceinline__ F2 mul2_nc(F2 a, F2 b) {
    F2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(0ull));
    return r;
}
// certified exact division of both halves by one constant (Div<DM_FAST> on pairs)
struct Div2 {
    F2 inv, inv_lo;
    __device__ __forceinline__ F2 operator()(F2 x) const { return fma2(x, inv, mul2(x, inv_lo)); }
};
template <int DM>
__device__ __forceinline__ Div2 make_div2(const Div<DM>& d) {
    Div2 r;
    r.inv = pk1(d.inv);
    r.inv_lo = pk1(d.inv_lo);
    return r;
}
// out = fma(v_se,se, fma(v_sw,sw, fma(v_ne,ne, v_nw*nw))) on pairs (see combine4 below)
__device__ __forceinline__ F2 combine4_2(F2 vnw, F2 vne, F2 vsw, F2 vse, F2 nw, F2 ne, F2 sw, F2 se) {
    return fma2(vse, se, fma2(vsw, sw, fma2(vne, ne, mul2(vnw, nw))));
}

// util.py:334-339  2*c/(size-1) - 1
template <class D>
__device__ __forceinline__ float norm_coord(float c, const D& d) { return sub(d(mul(2.0f, c)), 1.0f); }
// util.py:343-348  2*f/(size-1)
template <class D>
__device__ __forceinline__ float norm_flow(float f, const D& d) { return d(mul(2.0f, f)); }
// Fused forms for the certified division: 2*v/s and v/(s/2) are the same real quotient, 2*v and
// s/2 are exact, so RN(RN(2v)/s) == RN(v/(s/2)) (barring overflow of 2v, |v| > 1.7e38).  `dh2`
// must be a Div over s/2.
template <class D>
__device__ __forceinline__ float norm_flow_h(float f, const D& dhalf) { return dhalf(f); }
template <class D>
__device__ __forceinline__ float norm_coord_h(float c, const D& dhalf) { return sub(dhalf(c), 1.0f); }
// util.py:352-357  (f*(size-1))/2
__device__ __forceinline__ float denorm_flow(float f, float size_m1) { return mul(mul(f, size_m1), 0.5f); }

// ---- ATen grid_sampler_2d (bilinear, zeros padding, align_corners=True) -------------------
// Unnormalise + corner weights.  half_w = (W-1)/2 exactly representable for W < 2^24.
struct Taps {
    int x0, y0;          // north-west corner
    float nw, ne, sw, se;
    bool inx0, inx1, iny0, iny1;
};
__device__ __forceinline__ Taps make_taps(float gx, float gy, int W, int H, float half_w, float half_h) {
    Taps t;
    float ix = mul(add(gx, 1.0f), half_w);
    float iy = mul(add(gy, 1.0f), half_h);
    float xw = floorf(ix), yn = floorf(iy);
    float xe = add(xw, 1.0f), ys = add(yn, 1.0f);
    float w = sub(ix, xw), e = sub(xe, ix), n = sub(iy, yn), s = sub(ys, iy);
    t.nw = mul(s, e);
    t.ne = mul(s, w);
    t.sw = mul(n, e);
    t.se = mul(n, w);
    // float comparisons so that NaN / huge coordinates select no tap, like ATen's masks
    t.inx0 = (xw > -1.0f) && (xw < (float)W);
    t.inx1 = (xe > -1.0f) && (xe < (float)W);
    t.iny0 = (yn > -1.0f) && (yn < (float)H);
    t.iny1 = (ys > -1.0f) && (ys < (float)H);
    t.x0 = t.inx0 ? (int)xw : ((t.inx1) ? -1 : 0);
    t.y0 = t.iny0 ? (int)yn : ((t.iny1) ? -1 : 0);
    return t;
}
// Tap accumulation of ATen's grid_sampler_2d (CPU kernel, pinned bitwise; the native CUDA kernel
// has the same order — sm_100 SASS of torch 2.11: @!P0 FFMA acc=nw*v_nw+0; @!P1 FFMA ne; @!P2 sw;
// @!P3 se):  out = fma(v_se,se, fma(v_sw,sw, fma(v_ne,ne, v_nw*nw))).
// NOTE: with cuDNN enabled, torch dispatches bilinear/zeros/align_corners=True grid_sample on CUDA
// to cudnnSpatialTfSamplerForward instead, whose rounding differs (north-east tap first in the
// interior, another pattern at the frame border; measured on a B200) and is not reproduced here.
__device__ __forceinline__ float combine4(float vnw, float vne, float vsw, float vse, float nw, float ne, float sw, float se) {
    return fma_(vse, se, fma_(vsw, sw, fma_(vne, ne, mul(vnw, nw))));
}
__device__ __forceinline__ float combine(const Taps& t, float vnw, float vne, float vsw, float vse) {
    return combine4(vnw, vne, vsw, vse, t.nw, t.ne, t.sw, t.se);
}

// ---- ATen upsample_bilinear2d (align_corners=True) axis taps ------------------------------
struct AxisTap {
    int i0, i1;
    float l0, l1;
};
__device__ __forceinline__ AxisTap axis_tap(int d, float scale, int in) {
    AxisTap t;
    float s = mul(scale, (float)d);
    int i0 = (int)s;  // s >= 0: trunc == floor
    i0 = min(i0, in - 1);
    float l1 = sub(s, (float)i0);
    l1 = fminf(fmaxf(l1, 0.0f), 1.0f);
    t.i0 = i0;
    t.i1 = i0 + (i0 < in - 1 ? 1 : 0);
    t.l1 = l1;
    t.l0 = sub(1.0f, l1);
    return t;
}
// val = fma(l0y, fma(l0x,a, l1x*b), l1y*fma(l0x,c, l1x*d))
__device__ __forceinline__ float up_combine(const AxisTap& ty, const AxisTap& tx, float a, float b, float c, float d) {
    float top = fma_(tx.l0, a, mul(tx.l1, b));
    float bot = fma_(tx.l0, c, mul(tx.l1, d));
    return fma_(ty.l0, top, mul(ty.l1, bot));
}
static inline float up_scale(int in, int out) { return out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f; }

}  // namespace pp
