/*
 * pixpro_oracle.c — CPU restatement of the PixPro-with-OpticalFlow pixel-pretext hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pixpro-with-opticalflow_b200/)
 * may import, link or execute this file.  It is used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, as the checker and the CPU arm.
 *
 * Parity status: PINNED against the reference's own PyTorch implementation executed on
 * CPU in the build container (oracle/pin_against_reference.py imports /root/reference,
 * runs both on the same seeded inputs and writes tests/golden/).  The reference ships no
 * golden vectors or tests of its own, so outputs of the reference run on
 * CPU are the pin.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference).  Arithmetic that decides a boolean (positive masks, FB masks,
 * nearest-mask lookups) is restated op by op in IEEE fp32 with the same rounding points
 * as the reference's sequence of torch kernels; this file MUST be compiled with
 * -ffp-contract=off so that only the explicit fmaf() calls fuse.  Third-party arithmetic
 * on the path is PyTorch ATen (torch 2.11 here; reference pins torch 1.8.2):
 *   - F.grid_sample bilinear, align_corners=True, zeros padding:
 *       ix = (gx+1)*((W-1)/2); taps weighted by (x1-ix)*(y1-iy) etc. (one fmul each);
 *       out = fma(v_se,se, fma(v_sw,sw, fma(v_ne,ne, v_nw*nw)))   [verified bitwise vs
 *       torch CPU here; identical to the CUDA kernel's pattern, SURVEY.md A.1]
 *   - F.grid_sample nearest: nearbyint (ties to even)
 *   - F.interpolate bilinear align_corners=True:
 *       s = fl((in-1)/(out-1)) * dst; i0 = floor(s); l1 = s-i0; l0 = 1-l1;
 *       val = fma(l0y, fma(l0x,a, l1x*b), l1y*fma(l0x,c, l1x*d))   [verified bitwise]
 *   - tensor / python_scalar: true IEEE division on CPU (div_mode 0); torch's CUDA
 *     kernel multiplies by the fp32 reciprocal instead (div_mode 1).
 *     div_mode 1 reproduces the reference run through torch's NATIVE CUDA kernels (checked
 *     bitwise on a B200 with cuDNN disabled; cuDNN's own grid sampler rounds differently).
 * Reductions that only feed tolerance-checked floats (q·k logits, PPM contractions,
 * norms) are accumulated in double: the oracle is the more accurate side there.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/* helpers                                                                              */
/* ------------------------------------------------------------------------------------ */

/* tensor / python scalar.  div_mode 0: CPU torch (true division). 1: CUDA torch (x * fl(1/s)). */
static inline float div_scalar(float x, float s, int div_mode) {
    if (div_mode) {
        float inv = 1.0f / s;
        return x * inv;
    }
    return x / s;
}

/* util.py:334-339 normalize_coord: 2 * c / (size-1) - 1  (mul, div, sub; each rounded) */
static inline float norm_coord1(float c, int size, int div_mode) {
    float t = 2.0f * c;
    t = div_scalar(t, (float)(size - 1), div_mode);
    return t - 1.0f;
}

/* util.py:343-348 normalize_flow: 2 * f / (size-1) */
static inline float norm_flow1(float f, int size, int div_mode) {
    float t = 2.0f * f;
    return div_scalar(t, (float)(size - 1), div_mode);
}

/* util.py:352-357 denormalize_flow: (f * (size-1)) / 2 */
static inline float denorm_flow1(float f, int size) {
    float t = f * (float)(size - 1);
    return t / 2.0f; /* exact either way */
}

/* ATen grid_sampler_2d, bilinear, zeros padding, align_corners=True, C channels of one
 * sample.  in: [C,H,W] with channel stride cs; (gx,gy) normalised; out[c] written. */
/* tap accumulation order of ATen's CPU kernel (and of its native CUDA kernel: same order in the
 * sm_100 SASS).  cuDNN's sampler, which torch prefers on CUDA when cuDNN is enabled, rounds
 * differently and is not restated. */
static inline float tap_combine(float vnw, float vne, float vsw, float vse, float nw, float ne, float sw, float se,
                                int unused) {
    (void)unused;
    return fmaf(vse, se, fmaf(vsw, sw, fmaf(vne, ne, vnw * nw)));
}

static inline void grid_sample_bilinear_pt(const float* in, long cs, int C, int H, int W,
                                           float gx, float gy, float* out, int cuda_order) {
    float ix = (gx + 1.0f) * ((float)(W - 1) / 2.0f);
    float iy = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
    float xw = floorf(ix), yn = floorf(iy);
    float xe = xw + 1.0f, ys = yn + 1.0f;
    float w = ix - xw, e = xe - ix, n = iy - yn, s = ys - iy;
    float nw = s * e, ne = s * w, sw = n * e, se = n * w;
    /* NaN/inf coordinates: every comparison below is false -> all taps zero, like ATen */
    int inx0 = (xw > -1.0f) && (xw < (float)W);
    int inx1 = (xe > -1.0f) && (xe < (float)W);
    int iny0 = (yn > -1.0f) && (yn < (float)H);
    int iny1 = (ys > -1.0f) && (ys < (float)H);
    long x0 = inx0 ? (long)xw : 0, x1 = inx1 ? (long)xe : 0;
    long y0 = iny0 ? (long)yn : 0, y1 = iny1 ? (long)ys : 0;
    for (int c = 0; c < C; c++) {
        const float* p = in + (long)c * cs;
        float vnw = (inx0 && iny0) ? p[y0 * W + x0] : 0.0f;
        float vne = (inx1 && iny0) ? p[y0 * W + x1] : 0.0f;
        float vsw = (inx0 && iny1) ? p[y1 * W + x0] : 0.0f;
        float vse = (inx1 && iny1) ? p[y1 * W + x1] : 0.0f;
        out[c] = tap_combine(vnw, vne, vsw, vse, nw, ne, sw, se, cuda_order);
    }
}

/* ATen grid_sampler_2d nearest, zeros padding, align_corners=True on a u8 mask plane
 * (reference converts bool->float, samples, converts back: PixPro.py:65-70). */
static inline uint8_t grid_sample_nearest_mask(const uint8_t* m, int H, int W, float gx, float gy) {
    float ix = (gx + 1.0f) * ((float)(W - 1) / 2.0f);
    float iy = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
    float xr = nearbyintf(ix), yr = nearbyintf(iy);
    if (!(xr > -1.0f && xr < (float)W && yr > -1.0f && yr < (float)H)) return 0;
    return m[(long)yr * W + (long)xr] != 0;
}

/* ------------------------------------------------------------------------------------ */
/* a1  upflow8  — contrast/flow/utils/utils.py:87-89                                     */
/* ------------------------------------------------------------------------------------ */

typedef struct { int i0, i1; float l0, l1; } axis_tap;

static void make_axis_taps(int in, int out, axis_tap* t) {
    /* ATen area_pixel_compute_scale (align_corners): (in-1)/(out-1) in fp32 */
    float scale = out > 1 ? (float)(in - 1) / (float)(out - 1) : 0.0f;
    for (int d = 0; d < out; d++) {
        float s = scale * (float)d;
        int i0 = (int)floorf(s);
        if (i0 > in - 1) i0 = in - 1;
        float l1 = s - (float)i0;
        if (l1 < 0.0f) l1 = 0.0f;
        if (l1 > 1.0f) l1 = 1.0f;
        t[d].i0 = i0;
        t[d].i1 = i0 + (i0 < in - 1 ? 1 : 0);
        t[d].l1 = l1;
        t[d].l0 = 1.0f - l1;
    }
}

static inline float up_eval(const float* p, int w, const axis_tap* ty, const axis_tap* tx) {
    float a = p[(long)ty->i0 * w + tx->i0], b = p[(long)ty->i0 * w + tx->i1];
    float c = p[(long)ty->i1 * w + tx->i0], d = p[(long)ty->i1 * w + tx->i1];
    float top = fmaf(tx->l0, a, tx->l1 * b);
    float bot = fmaf(tx->l0, c, tx->l1 * d);
    return fmaf(ty->l0, top, ty->l1 * bot);
}

/* in: planes [NP,h,w] (NP = N*2) -> out [NP,8h,8w];  8 * interpolate(bilinear, align_corners) */
ORC_API void orc_upflow8(const float* in, long NP, int h, int w, float* out) {
    int H = 8 * h, W = 8 * w;
    axis_tap* ty = (axis_tap*)malloc(sizeof(axis_tap) * H);
    axis_tap* tx = (axis_tap*)malloc(sizeof(axis_tap) * W);
    make_axis_taps(h, H, ty);
    make_axis_taps(w, W, tx);
#pragma omp parallel for collapse(2) schedule(static)
    for (long p = 0; p < NP; p++)
        for (int Y = 0; Y < H; Y++) {
            const float* src = in + p * (long)h * w;
            float* dst = out + (p * H + Y) * (long)W;
            for (int X = 0; X < W; X++) dst[X] = 8.0f * up_eval(src, w, &ty[Y], &tx[X]);
        }
    free(ty);
    free(tx);
}

/* ------------------------------------------------------------------------------------ */
/* a2  normalize_coord / normalize_flow / denormalize_flow — contrast/util.py:334-357    */
/* ------------------------------------------------------------------------------------ */

/* kind: 0 normalize_coord, 1 normalize_flow, 2 denormalize_flow; x: [B,2,H,W] */
ORC_API void orc_normalize(const float* x, long B, int H, int W, int kind, int div_mode, float* out) {
#pragma omp parallel for schedule(static)
    for (long bc = 0; bc < B * 2; bc++) {
        int size = (bc & 1) ? H : W;
        const float* s = x + bc * (long)H * W;
        float* d = out + bc * (long)H * W;
        for (long i = 0; i < (long)H * W; i++) {
            if (kind == 0) d[i] = norm_coord1(s[i], size, div_mode);
            else if (kind == 1) d[i] = norm_flow1(s[i], size, div_mode);
            else d[i] = denorm_flow1(s[i], size);
        }
    }
}

/* generic F.grid_sample(bilinear, zeros, align_corners=True): in [N,C,H,W], grid [N,Ho,Wo,2] */
ORC_API void orc_grid_sample_bilinear(const float* in, long N, int C, int H, int W,
                                      const float* grid, int Ho, int Wo, float* out) {
#pragma omp parallel for schedule(static)
    for (long n = 0; n < N; n++) {
        float tmp[16];
        for (long i = 0; i < (long)Ho * Wo; i++) {
            const float* g = grid + (n * Ho * Wo + i) * 2;
            for (int c0 = 0; c0 < C; c0 += 16) {
                int cc = C - c0 < 16 ? C - c0 : 16;
                grid_sample_bilinear_pt(in + (n * C + c0) * (long)H * W, (long)H * W, cc, H, W, g[0], g[1], tmp, 0);
                for (int c = 0; c < cc; c++) out[(n * C + c0 + c) * (long)Ho * Wo + i] = tmp[c];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* a3  concat_flow — contrast/util.py:301-330                                            */
/*     flows: n chain links; link i, sample b is the [2,H,W] block at                     */
/*     flows + i*stride_n + b*stride_b  (so both [n,B,2,H,W] and the loader's             */
/*     [B,n,2,H,W] layouts are addressable).  out [B,2,H,W].                              */
/* ------------------------------------------------------------------------------------ */
ORC_API void orc_concat_flow(const float* flows, int n, long B, int H, int W,
                             long stride_n, long stride_b, int is_norm, int div_mode, float* out) {
    long HW = (long)H * W;
    if (n == 1) { /* util.py:303-308: clone (normalised if is_norm) */
#pragma omp parallel for schedule(static)
        for (long b = 0; b < B; b++)
            for (int c = 0; c < 2; c++) {
                const float* s = flows + b * stride_b + c * HW;
                float* d = out + (b * 2 + c) * HW;
                int size = c ? H : W;
                for (long i = 0; i < HW; i++) d[i] = is_norm ? norm_flow1(s[i], size, div_mode) : s[i];
            }
        return;
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < B; b++)
        for (int Y = 0; Y < H; Y++) {
            float* ox = out + (b * 2 + 0) * HW + (long)Y * W;
            float* oy = out + (b * 2 + 1) * HW + (long)Y * W;
            for (int X = 0; X < W; X++) {
                float c0x = (float)X, c0y = (float)Y; /* util.py:309-311 meshgrid (x,y) */
                float s[2];
                if (!is_norm) {
                    float cx = c0x, cy = c0y;
                    for (int i = 0; i < n; i++) { /* util.py:321-323 */
                        const float* f = flows + i * stride_n + b * stride_b;
                        float gx = norm_coord1(cx, W, div_mode), gy = norm_coord1(cy, H, div_mode);
                        grid_sample_bilinear_pt(f, HW, 2, H, W, gx, gy, s, div_mode);
                        cx = cx + s[0];
                        cy = cy + s[1];
                    }
                    ox[X] = cx - c0x; /* util.py:328 */
                    oy[X] = cy - c0y;
                } else {
                    /* util.py:316-319: sample normalize_flow(flow) at the running normalised
                     * coordinate.  Sampling the normalised field: each tap is normalised first. */
                    float n0x = norm_coord1(c0x, W, div_mode), n0y = norm_coord1(c0y, H, div_mode);
                    float cx = n0x, cy = n0y;
                    for (int i = 0; i < n; i++) {
                        const float* f = flows + i * stride_n + b * stride_b;
                        /* inline bilinear on normalised taps */
                        float ix = (cx + 1.0f) * ((float)(W - 1) / 2.0f);
                        float iy = (cy + 1.0f) * ((float)(H - 1) / 2.0f);
                        float xw = floorf(ix), yn = floorf(iy), xe = xw + 1.0f, ys = yn + 1.0f;
                        float w = ix - xw, e = xe - ix, nn = iy - yn, ss = ys - iy;
                        float nw = ss * e, ne = ss * w, sw = nn * e, se = nn * w;
                        int inx0 = (xw > -1.0f) && (xw < (float)W), inx1 = (xe > -1.0f) && (xe < (float)W);
                        int iny0 = (yn > -1.0f) && (yn < (float)H), iny1 = (ys > -1.0f) && (ys < (float)H);
                        long x0 = inx0 ? (long)xw : 0, x1 = inx1 ? (long)xe : 0;
                        long y0 = iny0 ? (long)yn : 0, y1 = iny1 ? (long)ys : 0;
                        for (int c = 0; c < 2; c++) {
                            const float* p = f + c * HW;
                            int size = c ? H : W;
                            float vnw = (inx0 && iny0) ? norm_flow1(p[y0 * W + x0], size, div_mode) : 0.0f;
                            float vne = (inx1 && iny0) ? norm_flow1(p[y0 * W + x1], size, div_mode) : 0.0f;
                            float vsw = (inx0 && iny1) ? norm_flow1(p[y1 * W + x0], size, div_mode) : 0.0f;
                            float vse = (inx1 && iny1) ? norm_flow1(p[y1 * W + x1], size, div_mode) : 0.0f;
                            s[c] = tap_combine(vnw, vne, vsw, vse, nw, ne, sw, se, div_mode);
                        }
                        cx = cx + s[0];
                        cy = cy + s[1];
                    }
                    ox[X] = cx - n0x; /* util.py:326 */
                    oy[X] = cy - n0y;
                }
            }
        }
}

/* ------------------------------------------------------------------------------------ */
/* a5  forward_backward_consistency — contrast/util.py:253-297                           */
/*     fwd,bwd [B,2,H,W] -> mask u8 [B,H,W]; optional cycle [B,2,H,W] (normalised units), */
/*     optional coords1_norm [B,2,H,W].  is_norm: inputs already normalised (:258-262).   */
/* ------------------------------------------------------------------------------------ */
ORC_API float orc_fb_alpha2_eff(double alpha_2, int H, int W) {
    /* util.py:289-291: h,w int64 tensors; sqrt(int64) -> fp32; .item() -> double; python / */
    float r = sqrtf((float)((long)H * H + (long)W * W));
    double a2 = alpha_2 / (double)r;
    return (float)a2; /* cast to fp32 when added to the fp32 tensor (:294) */
}

ORC_API void orc_fb_consistency(const float* fwd, const float* bwd, long B, int H, int W,
                                double alpha_1, double alpha_2, int is_norm, int div_mode,
                                uint8_t* mask, float* cycle, float* coords1_norm) {
    long HW = (long)H * W;
    float a1 = (float)alpha_1;
    float a2 = orc_fb_alpha2_eff(alpha_2, H, W);
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < B; b++)
        for (int Y = 0; Y < H; Y++) {
            const float* fx = fwd + (b * 2 + 0) * HW;
            const float* fy = fwd + (b * 2 + 1) * HW;
            const float* bw = bwd + (b * 2) * HW;
            for (int X = 0; X < W; X++) {
                long i = (long)Y * W + X;
                float fnx = is_norm ? fx[i] : norm_flow1(fx[i], W, div_mode); /* :264 */
                float fny = is_norm ? fy[i] : norm_flow1(fy[i], H, div_mode);
                float c1x = norm_coord1((float)X, W, div_mode) + fnx;         /* :271,275 */
                float c1y = norm_coord1((float)Y, H, div_mode) + fny;
                int inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);         /* :276 */
                /* :278 grid_sample of the NORMALISED bwd flow at c1 */
                float ix = (c1x + 1.0f) * ((float)(W - 1) / 2.0f);
                float iy = (c1y + 1.0f) * ((float)(H - 1) / 2.0f);
                float xw = floorf(ix), yn = floorf(iy), xe = xw + 1.0f, ys = yn + 1.0f;
                float w = ix - xw, e = xe - ix, nn = iy - yn, ss = ys - iy;
                float nw = ss * e, ne = ss * w, sw = nn * e, se = nn * w;
                int inx0 = (xw > -1.0f) && (xw < (float)W), inx1 = (xe > -1.0f) && (xe < (float)W);
                int iny0 = (yn > -1.0f) && (yn < (float)H), iny1 = (ys > -1.0f) && (ys < (float)H);
                long x0 = inx0 ? (long)xw : 0, x1 = inx1 ? (long)xe : 0;
                long y0 = iny0 ? (long)yn : 0, y1 = iny1 ? (long)ys : 0;
                float bi[2];
                for (int c = 0; c < 2; c++) {
                    const float* p = bw + c * HW;
                    int size = c ? H : W;
                    float vnw = (inx0 && iny0) ? p[y0 * W + x0] : 0.0f;
                    float vne = (inx1 && iny0) ? p[y0 * W + x1] : 0.0f;
                    float vsw = (inx0 && iny1) ? p[y1 * W + x0] : 0.0f;
                    float vse = (inx1 && iny1) ? p[y1 * W + x1] : 0.0f;
                    if (!is_norm) { /* zero taps stay zero under normalisation (2*0/s = 0) */
                        vnw = norm_flow1(vnw, size, div_mode);
                        vne = norm_flow1(vne, size, div_mode);
                        vsw = norm_flow1(vsw, size, div_mode);
                        vse = norm_flow1(vse, size, div_mode);
                    }
                    bi[c] = tap_combine(vnw, vne, vsw, vse, nw, ne, sw, se, div_mode);
                }
                float cyx = fnx + bi[0], cyy = fny + bi[1];                    /* :279 */
                float cyc2 = cyx * cyx + cyy * cyy;                            /* :293 */
                float f2 = fnx * fnx + fny * fny;
                float b2 = bi[0] * bi[0] + bi[1] * bi[1];
                float eps = a1 * (f2 + b2) + a2;                               /* :294 */
                int ok = inb && ((cyc2 - eps) <= 0.0f);                        /* :296 */
                mask[b * HW + i] = (uint8_t)ok;
                if (cycle) { cycle[(b * 2) * HW + i] = cyx; cycle[(b * 2 + 1) * HW + i] = cyy; }
                if (coords1_norm) { coords1_norm[(b * 2) * HW + i] = c1x; coords1_norm[(b * 2 + 1) * HW + i] = c1y; }
            }
        }
}

/* a11 calc_mask_ratio — contrast/util.py:361-366: mean over W then over H of !mask */
ORC_API void orc_calc_mask_ratio(const uint8_t* mask, long B, int H, int W, float* ratio) {
    for (long b = 0; b < B; b++) {
        double acc = 0;
        for (long i = 0; i < (long)H * W; i++) acc += mask[b * (long)H * W + i] ? 0.0 : 1.0;
        ratio[b] = (float)(acc / ((double)H * W));
    }
}

/* ------------------------------------------------------------------------------------ */
/* a7  add_optical_flow — contrast/models/PixPro.py:46-89                                */
/*     flow [B,2,Hin,Win]; x_grid,y_grid [B,P] pixel coords in the ORIGINAL frame;        */
/*     mask u8 [B,Hin,Win] or NULL.  -> out_x,out_y [B,P], mask_grid u8 [B,P] (or NULL)   */
/* ------------------------------------------------------------------------------------ */
ORC_API void orc_add_optical_flow(const float* flow, long B, int Hin, int Win,
                                  const float* x_grid, const float* y_grid, int P,
                                  int H_orig, int W_orig, const uint8_t* mask, int div_mode,
                                  float* out_x, float* out_y, uint8_t* mask_grid) {
    long HW = (long)Hin * Win;
    int diff = (Hin != H_orig) || (Win != W_orig);
    float rh = (float)((double)Hin / (double)H_orig), rw = (float)((double)Win / (double)W_orig);
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; b++)
        for (int p = 0; p < P; p++) {
            float xg = x_grid[b * P + p], yg = y_grid[b * P + p];
            /* :61-62  2 * (x / (W_orig-1)) - 1 */
            float gx = 2.0f * div_scalar(xg, (float)(W_orig - 1), div_mode) - 1.0f;
            float gy = 2.0f * div_scalar(yg, (float)(H_orig - 1), div_mode) - 1.0f;
            float fg[2];
            grid_sample_bilinear_pt(flow + b * 2 * HW, HW, 2, Hin, Win, gx, gy, fg, div_mode); /* :64 */
            if (mask_grid) mask_grid[b * P + p] = mask ? grid_sample_nearest_mask(mask + b * HW, Hin, Win, gx, gy) : 1;
            if (diff) { /* :76-80 */
                float ox = xg * rw + fg[0];
                float oy = yg * rh + fg[1];
                out_x[b * P + p] = div_scalar(ox, rw, div_mode);
                out_y[b * P + p] = div_scalar(oy, rh, div_mode);
            } else { /* :82-83 */
                out_x[b * P + p] = xg + fg[0];
                out_y[b * P + p] = yg + fg[1];
            }
        }
}

/* ------------------------------------------------------------------------------------ */
/* a8  regression_loss — contrast/models/PixPro.py:92-247                                */
/* ------------------------------------------------------------------------------------ */

/* Grid centres in original-frame pixels (:140-143, :168-175 / :192-199) and the bin
 * diagonal (:155).  coord: [B,10]; out cx,cy [B,P] (P=G*G, row-major y*G+x), diag [B]. */
ORC_API void orc_grid_centres(const float* coord, long B, int G, int H_orig, int W_orig, int div_mode,
                              float* cx, float* cy, float* diag) {
    int P = G * G;
    float wo = (float)(W_orig - 1), ho = (float)(H_orig - 1);
    for (long b = 0; b < B; b++) {
        const float* c = coord + b * 10;
        float bw = div_scalar(c[2] - c[0], (float)G, div_mode);
        float bh = div_scalar(c[3] - c[1], (float)G, div_mode);
        for (int y = 0; y < G; y++)
            for (int x = 0; x < G; x++) {
                float vx = ((float)x + 0.5f) * bw;
                vx = vx + c[0];
                float vy = ((float)y + 0.5f) * bh;
                vy = vy + c[1];
                cx[b * P + y * G + x] = vx * wo;
                cy[b * P + y * G + x] = vy * ho;
            }
        float dw = bw * wo, dh = bh * ho;
        float d2 = dw * dw + dh * dh;
        diag[b] = sqrtf(d2);
    }
}

/* Full regression_loss.
 *   q,k [B,C,P] (q = prediction with grad, k = detached key); coord_q/coord_k [B,10];
 *   flow [B,2,Hin,Win] or NULL (no-flow path :167-175); mask u8 [B,Hin,Win] or NULL.
 * Outputs (any may be NULL except loss): loss[1]; pos_num[B]; pos_mean[B];
 *   pos_mask u8 [B,P,P] (row i = query cell, col j = key cell); cqx,cqy,ckx,cky [B,P]
 *   (the flow-warped query centres and the key centres — the "correspondence");
 *   dq [B,C,P] = d loss / d q  (verified vs autograd). */
ORC_API void orc_regression_loss(const float* q, const float* k, long B, int C, int G,
                                 const float* coord_q, const float* coord_k,
                                 const float* flow, int Hin, int Win, const uint8_t* mask,
                                 int H_orig, int W_orig, double pos_ratio, int div_mode,
                                 float* loss, float* pos_num, float* pos_mean, uint8_t* pos_mask,
                                 float* cqx_o, float* cqy_o, float* ckx_o, float* cky_o, float* dq) {
    int P = G * G;
    float* qx = (float*)malloc(sizeof(float) * B * P);
    float* qy = (float*)malloc(sizeof(float) * B * P);
    float* kx = (float*)malloc(sizeof(float) * B * P);
    float* ky = (float*)malloc(sizeof(float) * B * P);
    float* qd = (float*)malloc(sizeof(float) * B);
    float* kd = (float*)malloc(sizeof(float) * B);
    uint8_t* mg = (uint8_t*)malloc((size_t)B * P);
    uint8_t* pm = pos_mask ? pos_mask : (uint8_t*)malloc((size_t)B * P * P);
    double* lossb = (double*)malloc(sizeof(double) * B);
    orc_grid_centres(coord_q, B, G, H_orig, W_orig, div_mode, qx, qy, qd);
    orc_grid_centres(coord_k, B, G, H_orig, W_orig, div_mode, kx, ky, kd);
    memset(mg, 1, (size_t)B * P);
    if (flow) { /* :200 */
        float* wx = (float*)malloc(sizeof(float) * B * P);
        float* wy = (float*)malloc(sizeof(float) * B * P);
        orc_add_optical_flow(flow, B, Hin, Win, qx, qy, P, H_orig, W_orig, mask, div_mode, wx, wy, mask ? mg : NULL);
        memcpy(qx, wx, sizeof(float) * B * P);
        memcpy(qy, wy, sizeof(float) * B * P);
        free(wx);
        free(wy);
    }
    float pr = (float)pos_ratio;
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; b++) {
        float md = qd[b] > kd[b] ? qd[b] : kd[b]; /* :157 */
        long cnt = 0;
        double acc = 0.0;
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) {
                float dx = qx[b * P + i] - kx[b * P + j];
                float dy = qy[b * P + i] - ky[b * P + j];
                float d2 = dx * dx + dy * dy;
                float d = sqrtf(d2) / md;            /* :217-218 */
                int pos = (d < pr) && mg[b * P + i];  /* :219-222 */
                pm[(b * P + i) * (long)P + j] = (uint8_t)pos;
                if (pos) {
                    cnt++;
                    double dot = 0.0; /* :239 logit[i][j] = sum_c q[c][i] k[c][j] */
                    for (int c = 0; c < C; c++) dot += (double)q[(b * C + c) * (long)P + i] * (double)k[(b * C + c) * (long)P + j];
                    acc += dot;
                }
            }
        float den = (float)cnt + 1e-6f;              /* :241 fp32 denominator */
        lossb[b] = acc / (double)den;
        if (pos_num) pos_num[b] = (float)cnt;
        if (pos_mean) pos_mean[b] = (float)((double)cnt / ((double)P * P));
        if (dq) { /* dq[b,:,i] = -2/B * sum_j pos[i][j] k[b,:,j] / den */
            double sc = -2.0 / (double)B / (double)den;
            for (int c = 0; c < C; c++)
                for (int i = 0; i < P; i++) {
                    double s = 0.0;
                    for (int j = 0; j < P; j++)
                        if (pm[(b * P + i) * (long)P + j]) s += (double)k[(b * C + c) * (long)P + j];
                    dq[(b * C + c) * (long)P + i] = (float)(s * sc);
                }
        }
    }
    double tot = 0.0;
    for (long b = 0; b < B; b++) tot += lossb[b];
    loss[0] = (float)(-2.0 * tot / (double)B); /* :247 */
    if (cqx_o) memcpy(cqx_o, qx, sizeof(float) * B * P);
    if (cqy_o) memcpy(cqy_o, qy, sizeof(float) * B * P);
    if (ckx_o) memcpy(ckx_o, kx, sizeof(float) * B * P);
    if (cky_o) memcpy(cky_o, ky, sizeof(float) * B * P);
    free(qx); free(qy); free(kx); free(ky); free(qd); free(kd); free(mg); free(lossb);
    if (!pos_mask) free(pm);
}

/* ------------------------------------------------------------------------------------ */
/* a9  featprop (PPM) + the caller's F.normalize — contrast/models/PixPro.py:339-363,380 */
/*     feat [B,C,P] (projector output), val [B,C,P] (value_transform(feat), computed by   */
/*     the caller: Identity / 1x1 conv / MLP2d, :297-304).                                */
/*     out = normalize_c( normalize_c(val) · (clamp(x̂ᵀx̂, min=cv) [+1e-6 if p<1])^p ᵀ )    */
/*     if final_norm, else the un-normalised propagation result (featprop's own return).  */
/* ------------------------------------------------------------------------------------ */
static void normalize_cols(const float* u, int C, int P, double* uh, double* nrm) {
    for (int i = 0; i < P; i++) {
        double s = 0.0;
        for (int c = 0; c < C; c++) s += (double)u[(long)c * P + i] * (double)u[(long)c * P + i];
        double n = sqrt(s);
        if (n < 1e-12) n = 1e-12; /* F.normalize eps */
        nrm[i] = n;
        for (int c = 0; c < C; c++) uh[(long)c * P + i] = (double)u[(long)c * P + i] / n;
    }
}

ORC_API void orc_featprop(const float* feat, const float* val, long B, int C, int P,
                          double gamma, double clamp_value, int final_norm, float* out) {
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < B; b++) {
        double* xh = (double*)malloc(sizeof(double) * C * P);
        double* vh = (double*)malloc(sizeof(double) * C * P);
        double* nx = (double*)malloc(sizeof(double) * P);
        double* nv = (double*)malloc(sizeof(double) * P);
        double* A = (double*)malloc(sizeof(double) * P * P);
        double* y = (double*)malloc(sizeof(double) * C * P);
        normalize_cols(feat + b * (long)C * P, C, P, xh, nx);
        normalize_cols(val + b * (long)C * P, C, P, vh, nv);
        for (int i = 0; i < P; i++)
            for (int j = i; j < P; j++) {
                double s = 0.0;
                for (int c = 0; c < C; c++) s += xh[(long)c * P + i] * xh[(long)c * P + j];
                if (s < clamp_value) s = clamp_value;   /* :355 */
                if (gamma < 1.0) s += 1e-6;             /* :356-357 */
                s = (gamma == 2.0) ? s * s : (gamma == 1.0 ? s : pow(s, gamma)); /* :358 */
                A[(long)i * P + j] = s;
                A[(long)j * P + i] = s;
            }
        /* :361 out[c][i] = sum_j vh[c][j] * A[i][j] */
        for (int c = 0; c < C; c++)
            for (int i = 0; i < P; i++) {
                double s = 0.0;
                for (int j = 0; j < P; j++) s += vh[(long)c * P + j] * A[(long)i * P + j];
                y[(long)c * P + i] = s;
            }
        float* o = out + b * (long)C * P;
        if (final_norm) { /* :380 */
            for (int i = 0; i < P; i++) {
                double s = 0.0;
                for (int c = 0; c < C; c++) s += y[(long)c * P + i] * y[(long)c * P + i];
                double n = sqrt(s);
                if (n < 1e-12) n = 1e-12;
                for (int c = 0; c < C; c++) o[(long)c * P + i] = (float)(y[(long)c * P + i] / n);
            }
        } else {
            for (long t = 0; t < (long)C * P; t++) o[t] = (float)y[t];
        }
        free(xh); free(vh); free(nx); free(nv); free(A); free(y);
    }
}

/* Backward of the above w.r.t. feat (through the similarity only) and val.
 *   g [B,C,P] = dL/d out.  d_feat_sim [B,C,P]: gradient reaching feat through x̂ᵀx̂;
 *   d_val [B,C,P]: gradient w.r.t. val (caller back-propagates it through
 *   value_transform and adds to d_feat_sim).  Formulas: SURVEY.md §8 a9, restating
 *   autograd of PixPro.py:343-363,380; validated against torch autograd of the reference
 *   in oracle/pin_against_reference.py. */
ORC_API void orc_featprop_bwd(const float* feat, const float* val, const float* g, long B, int C, int P,
                              double gamma, double clamp_value, int final_norm,
                              float* d_feat_sim, float* d_val) {
#pragma omp parallel for schedule(dynamic)
    for (long b = 0; b < B; b++) {
        long CP = (long)C * P;
        double* xh = (double*)malloc(sizeof(double) * CP);
        double* vh = (double*)malloc(sizeof(double) * CP);
        double* nx = (double*)malloc(sizeof(double) * P);
        double* nv = (double*)malloc(sizeof(double) * P);
        double* S = (double*)malloc(sizeof(double) * P * P);
        double* A = (double*)malloc(sizeof(double) * P * P);
        double* y = (double*)malloc(sizeof(double) * CP);
        double* gy = (double*)malloc(sizeof(double) * CP);
        double* gA = (double*)malloc(sizeof(double) * P * P);
        double* gS = (double*)malloc(sizeof(double) * P * P);
        double* gvh = (double*)malloc(sizeof(double) * CP);
        double* gxh = (double*)malloc(sizeof(double) * CP);
        const float* gb = g + b * CP;
        normalize_cols(feat + b * CP, C, P, xh, nx);
        normalize_cols(val + b * CP, C, P, vh, nv);
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) {
                double s = 0.0;
                for (int c = 0; c < C; c++) s += xh[(long)c * P + i] * xh[(long)c * P + j];
                S[(long)i * P + j] = s;
                double a = s < clamp_value ? clamp_value : s;
                if (gamma < 1.0) a += 1e-6;
                A[(long)i * P + j] = (gamma == 2.0) ? a * a : (gamma == 1.0 ? a : pow(a, gamma));
            }
        for (int c = 0; c < C; c++)
            for (int i = 0; i < P; i++) {
                double s = 0.0;
                for (int j = 0; j < P; j++) s += vh[(long)c * P + j] * A[(long)i * P + j];
                y[(long)c * P + i] = s;
            }
        /* through the final normalize: gy = (g - ŷ (g·ŷ)) / ||y|| */
        for (int i = 0; i < P; i++) {
            if (final_norm) {
                double s = 0.0, d = 0.0;
                for (int c = 0; c < C; c++) s += y[(long)c * P + i] * y[(long)c * P + i];
                double n = sqrt(s);
                if (n < 1e-12) n = 1e-12;
                for (int c = 0; c < C; c++) d += (double)gb[(long)c * P + i] * (y[(long)c * P + i] / n);
                for (int c = 0; c < C; c++) gy[(long)c * P + i] = ((double)gb[(long)c * P + i] - (y[(long)c * P + i] / n) * d) / n;
            } else {
                for (int c = 0; c < C; c++) gy[(long)c * P + i] = (double)gb[(long)c * P + i];
            }
        }
        /* y[c][i] = sum_j vh[c][j] A[i][j]  =>  gA[i][j] = sum_c gy[c][i] vh[c][j];  gvh[c][j] = sum_i gy[c][i] A[i][j] */
        for (int i = 0; i < P; i++)
            for (int j = 0; j < P; j++) {
                double s = 0.0;
                for (int c = 0; c < C; c++) s += gy[(long)c * P + i] * vh[(long)c * P + j];
                gA[(long)i * P + j] = s;
            }
        for (int c = 0; c < C; c++)
            for (int j = 0; j < P; j++) {
                double s = 0.0;
                for (int i = 0; i < P; i++) s += gy[(long)c * P + i] * A[(long)i * P + j];
                gvh[(long)c * P + j] = s;
            }
        /* A = (clamp(S,min=cv) [+1e-6])^p ; torch clamp backward passes where S >= cv */
        for (long t = 0; t < (long)P * P; t++) {
            double s = S[t];
            double a = s < clamp_value ? clamp_value : s;
            if (gamma < 1.0) a += 1e-6;
            double dA = (gamma == 2.0) ? 2.0 * a : (gamma == 1.0 ? 1.0 : gamma * pow(a, gamma - 1.0));
            gS[t] = (s >= clamp_value) ? gA[t] * dA : 0.0;
        }
        /* S = x̂ᵀx̂  =>  gx̂[c][i] = sum_j (gS[i][j] + gS[j][i]) x̂[c][j] */
        for (int c = 0; c < C; c++)
            for (int i = 0; i < P; i++) {
                double s = 0.0;
                for (int j = 0; j < P; j++) s += (gS[(long)i * P + j] + gS[(long)j * P + i]) * xh[(long)c * P + j];
                gxh[(long)c * P + i] = s;
            }
        /* through the two input normalisations */
        for (int i = 0; i < P; i++) {
            double dx = 0.0, dv = 0.0;
            for (int c = 0; c < C; c++) {
                dx += gxh[(long)c * P + i] * xh[(long)c * P + i];
                dv += gvh[(long)c * P + i] * vh[(long)c * P + i];
            }
            for (int c = 0; c < C; c++) {
                d_feat_sim[b * CP + (long)c * P + i] = (float)((gxh[(long)c * P + i] - xh[(long)c * P + i] * dx) / nx[i]);
                d_val[b * CP + (long)c * P + i] = (float)((gvh[(long)c * P + i] - vh[(long)c * P + i] * dv) / nv[i]);
            }
        }
        free(xh); free(vh); free(nx); free(nv); free(S); free(A); free(y); free(gy); free(gA); free(gS); free(gvh); free(gxh);
    }
}

/* ------------------------------------------------------------------------------------ */
/* a6  flow stage of apply_optical_flow (use_flow_file, not use_flow_frames)             */
/*     — contrast/util.py:175-248.  lo_fwd/lo_bwd: loader layout [B,n,2,h,w]              */
/*     (contrast/data/dataset.py:485-495).  flow_up: x8 upsample first (:185-191).        */
/*     Outputs: flow_fwd/flow_bwd [B,2,H,W]; mask_fwd/mask_bwd u8 [B,H,W] (NULL if        */
/*     alpha1/alpha2 are not set).  flow_cat_norm is restated through is_norm.            */
/* ------------------------------------------------------------------------------------ */
ORC_API void orc_flow_stage(const float* lo_fwd, const float* lo_bwd, long B, int n, int h, int w,
                            int flow_up, int use_mask, double alpha_1, double alpha_2,
                            int is_norm, int div_mode,
                            float* flow_fwd, float* flow_bwd, uint8_t* mask_fwd, uint8_t* mask_bwd) {
    int H = flow_up ? 8 * h : h, W = flow_up ? 8 * w : w;
    long HW = (long)H * W;
    const float* srcs[2] = { lo_fwd, lo_bwd };
    float* dsts[2] = { flow_fwd, flow_bwd };
    for (int d = 0; d < 2; d++) {
        const float* links = srcs[d];
        float* up = NULL;
        if (flow_up) {
            up = (float*)malloc(sizeof(float) * B * n * 2 * HW);
            orc_upflow8(srcs[d], B * n * 2, h, w, up);
            links = up;
        }
        /* layout [B,n,2,H,W]: link stride 2*HW, sample stride n*2*HW */
        orc_concat_flow(links, n, B, H, W, 2 * HW, (long)n * 2 * HW, is_norm, div_mode, dsts[d]);
        free(up);
    }
    if (use_mask) { /* :211-222 */
        orc_fb_consistency(flow_fwd, flow_bwd, B, H, W, alpha_1, alpha_2, is_norm, div_mode, mask_fwd, NULL, NULL);
        orc_fb_consistency(flow_bwd, flow_fwd, B, H, W, alpha_1, alpha_2, is_norm, div_mode, mask_bwd, NULL, NULL);
    }
    if (is_norm) { /* :229-231 */
        for (int d = 0; d < 2; d++) orc_normalize(dsts[d], B, H, W, 2, div_mode, dsts[d]);
    }
}

/* ------------------------------------------------------------------------------------ */
/* sparse correspondence — what regression_loss consumes of apply_optical_flow           */
/* (contrast/util.py:175-248) through add_optical_flow (contrast/models/PixPro.py:46-89),*/
/* restated WITHOUT the dense tensors: the composite flow (util.py:185-191,301-330) is   */
/* evaluated at single integer pixels, the FB test (util.py:253-297) at single pixels.   */
/* Every op of the dense stage is point-wise, so this must equal sampling the dense      */
/* outputs bit for bit — tests/test_oracle_golden.py checks it against orc_flow_stage +  */
/* orc_add_optical_flow and against the reference's own outputs in tests/golden.         */
/* ------------------------------------------------------------------------------------ */
typedef struct {
    const float* links; /* link i at links + i*stride_n: [2,h,w] (up) or [2,H,W] */
    int n, up, h, w, H, W, div_mode;
    long stride_n;
    const axis_tap *ty, *tx; /* up-sampling taps per full-res row / column (up only) */
} comp_field;

/* value of link i at integer pixel (Y,X): 8 * bilinear x8 (flow/utils/utils.py:87-89) or the dense value */
static inline void comp_link_value(const comp_field* f, int i, int Y, int X, float v[2]) {
    const float* p = f->links + i * f->stride_n;
    if (f->up) {
        v[0] = 8.0f * up_eval(p, f->w, &f->ty[Y], &f->tx[X]);
        v[1] = 8.0f * up_eval(p + (long)f->h * f->w, f->w, &f->ty[Y], &f->tx[X]);
    } else {
        v[0] = p[(long)Y * f->W + X];
        v[1] = p[(long)f->H * f->W + (long)Y * f->W + X];
    }
}

typedef void (*pixel_fn)(const comp_field*, int, int, int, float*);

/* F.grid_sample(bilinear, zeros, align_corners=True) of a 2-channel field given by `value` at normalised (gx,gy);
 * norm_taps: the sampled field is normalize_flow(field) (util.py:264-265,278) */
static inline void comp_grid_sample(const comp_field* f, pixel_fn value, int arg, float gx, float gy, int norm_taps, float out[2]) {
    int H = f->H, W = f->W;
    float ix = (gx + 1.0f) * ((float)(W - 1) / 2.0f);
    float iy = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
    float xw = floorf(ix), yn = floorf(iy), xe = xw + 1.0f, ys = yn + 1.0f;
    float w = ix - xw, e = xe - ix, nn = iy - yn, ss = ys - iy;
    float nw = ss * e, ne = ss * w, sw = nn * e, se = nn * w;
    int inx0 = (xw > -1.0f) && (xw < (float)W), inx1 = (xe > -1.0f) && (xe < (float)W);
    int iny0 = (yn > -1.0f) && (yn < (float)H), iny1 = (ys > -1.0f) && (ys < (float)H);
    int x0 = inx0 ? (int)xw : 0, x1 = inx1 ? (int)xe : 0, y0 = iny0 ? (int)yn : 0, y1 = iny1 ? (int)ys : 0;
    float t[4][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}};
    if (inx0 && iny0) value(f, arg, y0, x0, t[0]);
    if (inx1 && iny0) value(f, arg, y0, x1, t[1]);
    if (inx0 && iny1) value(f, arg, y1, x0, t[2]);
    if (inx1 && iny1) value(f, arg, y1, x1, t[3]);
    for (int c = 0; c < 2; c++) {
        float v[4];
        for (int k = 0; k < 4; k++) v[k] = norm_taps ? norm_flow1(t[k][c], c ? H : W, f->div_mode) : t[k][c];
        out[c] = tap_combine(v[0], v[1], v[2], v[3], nw, ne, sw, se, 0);
    }
}

/* composite flow at integer pixel (Y,X): util.py:303-308 (n == 1: the link itself) / :309-328 (chain) */
static void comp_value(const comp_field* f, int unused, int Y, int X, float v[2]) {
    (void)unused;
    if (f->n == 1) {
        comp_link_value(f, 0, Y, X, v);
        return;
    }
    float c0x = (float)X, c0y = (float)Y, cx = c0x, cy = c0y;
    for (int i = 0; i < f->n; i++) { /* util.py:321-323 */
        float gx = norm_coord1(cx, f->W, f->div_mode), gy = norm_coord1(cy, f->H, f->div_mode), s[2];
        comp_grid_sample(f, comp_link_value, i, gx, gy, 0, s);
        cx = cx + s[0];
        cy = cy + s[1];
    }
    v[0] = cx - c0x; /* util.py:328 */
    v[1] = cy - c0y;
}

/* lo_fwd/lo_bwd: loader layout [B,n,2,h,w] (dense links if !flow_up).  coord_fwd [B,10]: descriptors whose centres are
 * warped by the forward composite (mask = FB(fwd,bwd)); coord_bwd: by the backward one.  Either (coord, warped) pair may
 * be NULL.  warped_* [3,B,P]: out_x, out_y (PixPro.py:76-83), mask_grid as 0/1 (PixPro.py:65-70; ones if !use_mask). */
ORC_API void orc_sparse_corr(const float* lo_fwd, const float* lo_bwd, long B, int n, int h, int w, int flow_up, int use_mask,
                             double alpha_1, double alpha_2, const float* coord_fwd, const float* coord_bwd, int G,
                             int H_orig, int W_orig, int div_mode, float* warped_fwd, float* warped_bwd) {
    int H = flow_up ? 8 * h : h, W = flow_up ? 8 * w : w, P = G * G;
    axis_tap* ty = (axis_tap*)malloc(sizeof(axis_tap) * H);
    axis_tap* tx = (axis_tap*)malloc(sizeof(axis_tap) * W);
    if (flow_up) {
        make_axis_taps(h, H, ty);
        make_axis_taps(w, W, tx);
    }
    long link = 2L * h * w;
    float a1 = (float)alpha_1, a2 = orc_fb_alpha2_eff(alpha_2, H, W);
    int diff = (H != H_orig) || (W != W_orig);
    float rh = (float)((double)H / (double)H_orig), rw = (float)((double)W / (double)W_orig);
    for (int d = 0; d < 2; d++) {
        const float* coord = d ? coord_bwd : coord_fwd;
        float* out = d ? warped_bwd : warped_fwd;
        if (!coord || !out) continue;
        float* cx = (float*)malloc(sizeof(float) * B * P);
        float* cy = (float*)malloc(sizeof(float) * B * P);
        float* dg = (float*)malloc(sizeof(float) * B);
        orc_grid_centres(coord, B, G, H_orig, W_orig, div_mode, cx, cy, dg);
#pragma omp parallel for collapse(2) schedule(dynamic, 8)
        for (long b = 0; b < B; b++)
            for (int p = 0; p < P; p++) {
                comp_field f = {(d ? lo_bwd : lo_fwd) + b * n * link, n, flow_up, h, w, H, W, div_mode, link, ty, tx};
                comp_field g = {(d ? lo_fwd : lo_bwd) + b * n * link, n, flow_up, h, w, H, W, div_mode, link, ty, tx};
                float xg = cx[b * P + p], yg = cy[b * P + p];
                float gx = 2.0f * div_scalar(xg, (float)(W_orig - 1), div_mode) - 1.0f; /* PixPro.py:61-62 */
                float gy = 2.0f * div_scalar(yg, (float)(H_orig - 1), div_mode) - 1.0f;
                float fg[2];
                comp_grid_sample(&f, comp_value, 0, gx, gy, 0, fg); /* :64 */
                float mg = 1.0f;
                if (use_mask) { /* :65-70 nearest lookup of the FB mask, evaluated at that pixel only */
                    float ix = (gx + 1.0f) * ((float)(W - 1) / 2.0f), iy = (gy + 1.0f) * ((float)(H - 1) / 2.0f);
                    float xr = nearbyintf(ix), yr = nearbyintf(iy);
                    mg = 0.0f;
                    if (xr > -1.0f && xr < (float)W && yr > -1.0f && yr < (float)H) {
                        int X = (int)xr, Y = (int)yr;
                        float fv[2], bi[2];
                        comp_value(&f, 0, Y, X, fv);
                        float fnx = norm_flow1(fv[0], W, div_mode), fny = norm_flow1(fv[1], H, div_mode); /* util.py:264 */
                        float c1x = norm_coord1((float)X, W, div_mode) + fnx;                             /* :271,275 */
                        float c1y = norm_coord1((float)Y, H, div_mode) + fny;
                        int inb = (fabsf(c1x) < 1.0f) && (fabsf(c1y) < 1.0f);                             /* :276 */
                        comp_grid_sample(&g, comp_value, 0, c1x, c1y, 1, bi);                             /* :278 */
                        float cyx = fnx + bi[0], cyy = fny + bi[1];                                       /* :279 */
                        float cyc2 = cyx * cyx + cyy * cyy;                                               /* :293 */
                        float f2 = fnx * fnx + fny * fny, b2 = bi[0] * bi[0] + bi[1] * bi[1];
                        float eps = a1 * (f2 + b2) + a2;                                                  /* :294 */
                        mg = (inb && ((cyc2 - eps) <= 0.0f)) ? 1.0f : 0.0f;                               /* :296 */
                    }
                }
                float ox, oy;
                if (diff) { /* PixPro.py:76-80 */
                    ox = div_scalar(xg * rw + fg[0], rw, div_mode);
                    oy = div_scalar(yg * rh + fg[1], rh, div_mode);
                } else { /* :82-83 */
                    ox = xg + fg[0];
                    oy = yg + fg[1];
                }
                out[b * P + p] = ox;
                out[B * P + b * P + p] = oy;
                out[2 * B * P + b * P + p] = mg;
            }
        free(cx);
        free(cy);
        free(dg);
    }
    free(ty);
    free(tx);
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * optimizer-side loops (test oracle for pp_ema_update / pp_lars_sgd_step).
 * ------------------------------------------------------------------------------------------ */

/* contrast/models/PixPro.py:330-331  param_k = param_k * m + param_q * (1 - m)
 * (tensor * python float: the scalar is rounded to fp32; three separately rounded ops). */
ORC_API void orc_ema_update(float* k, const float* q, long n, double m, double one_minus_m) {
    const float mf = (float)m, omf = (float)one_minus_m;
    for (long i = 0; i < n; i++) {
        volatile float a = k[i] * mf;
        volatile float b = q[i] * omf;
        k[i] = a + b;
    }
}

/* One tensor of LARS.step() (contrast/lars.py:109-152) around torch.optim.SGD (no nesterov):
 *   g' = g.add(p, alpha=wd)            -> fmaf(wd, p, g)        (ATen add-with-alpha fuses)
 *   a  = trust * |p| / (|g'| + eps)    fp32 ops, if lars and both norms > 0, else 1
 *   g''= g' * a                        (only when lars)
 *   buf = g'' (first) | buf.mul_(mom).add_(g'', alpha=1-damp) ;  p.add_(buf, alpha=-lr)
 * Norms are accumulated in double (torch's fp32 reduction order is not reproduced: tolerance).
 * Returns the adaptive rate. */
ORC_API float orc_lars_sgd_step(float* p, const float* g, float* buf, long n, double wd, double lr, double mom, double damp,
                        int lars, int first, double trust, double eps) {
    const float wdf = (float)wd, nlr = -(float)lr, momf = (float)mom, omd = 1.0f - (float)damp;
    float a = 1.0f;
    if (lars) {
        double sp = 0.0, sg = 0.0;
        for (long i = 0; i < n; i++) {
            const float gv = wdf > 0.0f ? fmaf(wdf, p[i], g[i]) : g[i];
            sp += (double)p[i] * p[i];
            sg += (double)gv * gv;
        }
        const float pn = (float)sqrt(sp), gn = (float)sqrt(sg);
        if (pn > 0.0f && gn > 0.0f) {
            volatile float num = (float)trust * pn;
            volatile float den = gn + (float)eps;
            a = num / den;
        }
    }
    for (long i = 0; i < n; i++) {
        float gv = wdf > 0.0f ? fmaf(wdf, p[i], g[i]) : g[i];
        if (lars) { volatile float t = gv * a; gv = t; }
        if (momf != 0.0f) {
            float b;
            if (first) b = gv;
            else { volatile float t = buf[i] * momf; b = fmaf(omd, gv, t); }
            buf[i] = b;
            gv = b;
        }
        p[i] = fmaf(nlr, gv, p[i]);
    }
    return a;
}

/* ------------------------------------------------------------------------------------ */
/* SURVEY 8(f) rank 4 — RAFT correlation volume, pyramid, lookup                         */
/* contrast/flow/corr.py:12-60, contrast/flow/utils/utils.py:64-78                        */
/* ------------------------------------------------------------------------------------ */

/* corr.py:52-60: corr[b,i,j] = sum_d f1[b,d,i] f2[b,d,j] / sqrt(D).  torch.matmul accumulates in fp32 in an
 * unspecified order; the sum is formed in double here (tolerance oracle), the division is the reference's fp32
 * tensor / tensor division. */
ORC_API void orc_corr_volume(const float* f1, const float* f2, long B, int D, int h, int w, float* out) {
    const long P = (long)h * w;
    const float s = sqrtf((float)D);
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < B; b++)
        for (long i = 0; i < P; i++) {
            const float* a = f1 + b * D * P + i;
            const float* c = f2 + b * D * P;
            float* o = out + (b * P + i) * P;
            for (long j = 0; j < P; j++) {
                double acc = 0.0;
                for (int d = 0; d < D; d++) acc += (double)a[d * P] * (double)c[d * P + j];
                o[j] = (float)acc / s;
            }
        }
}

/* corr.py:26-28: F.avg_pool2d(corr, 2, stride=2): window summed row-major, divided by 4 (exact). */
ORC_API void orc_corr_pool(const float* in, long planes, int h, int w, float* out) {
    const int ho = h / 2, wo = w / 2;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < planes; p++)
        for (int y = 0; y < ho; y++)
            for (int x = 0; x < wo; x++) {
                const float* q = in + (p * h + 2 * y) * (long)w + 2 * x;
                float s = q[0] + q[1];
                s = s + q[w];
                s = s + q[w + 1];
                out[(p * ho + y) * (long)wo + x] = s * 0.25f;
            }
}

/* corr.py:30-50 + utils.py:64-78.  levels[l]: [B*P, h>>l, w>>l]; coords [B,2,h,w]; out [B, L*K*K, h, w].
 * delta = stack(meshgrid(dy, dx), -1) is added to (x, y): window index (i, j) shifts x by d[i] and y by d[j]. */
ORC_API void orc_corr_lookup(const float* const* levels, int L, const float* coords, long B, int h, int w, int r, int div_mode,
                             float* out) {
    const long P = (long)h * w;
    const int K = 2 * r + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (long b = 0; b < B; b++)
        for (long p = 0; p < P; p++) {
            const float cx = coords[(b * 2) * P + p], cy = coords[(b * 2 + 1) * P + p];
            int H = h, W = w;
            for (int l = 0; l < L; l++) {
                const float sc = (float)(1 << l);
                const float ccx = cx / sc, ccy = cy / sc;
                const float* plane = levels[l] + (b * P + p) * (long)H * W;
                for (int i = 0; i < K; i++)
                    for (int j = 0; j < K; j++) {
                        const float x = ccx + (float)(i - r), y = ccy + (float)(j - r);
                        const float gx = norm_coord1(x, W, div_mode), gy = norm_coord1(y, H, div_mode);
                        float v;
                        grid_sample_bilinear_pt(plane, 0, 1, H, W, gx, gy, &v, 0);
                        out[(b * (long)(L * K * K) + (long)l * K * K + i * K + j) * P + p] = v;
                    }
                H /= 2; W /= 2;
            }
        }
}
