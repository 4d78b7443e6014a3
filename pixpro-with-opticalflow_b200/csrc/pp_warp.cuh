// pp_warp.cuh — constants of add_optical_flow (contrast/models/PixPro.py:46-89), shared by the
// dense-flow loss kernels (pp_loss.cu) and the sparse correspondence kernel (pp_flow.cu).
#pragma once
#include "pp_common.cuh"

namespace pp {

struct WarpArgs {
    int Hin, Win;
    float half_w, half_h;       // (Win-1)/2, (Hin-1)/2
    ScalarDiv dwo, dho;         // / (W_orig-1), / (H_orig-1)
    ScalarDiv drw, drh;         // / ratio_w, / ratio_h
    float rw, rh;               // ratio_w = Win/W_orig, ratio_h = Hin/H_orig  (fp32 of the python double)
    int diff;                   // flow resolution != original resolution
};

static inline WarpArgs make_warp_args(int Hin, int Win, int H_orig, int W_orig, int div_mode) {
    WarpArgs a;
    a.Hin = Hin; a.Win = Win;
    a.half_w = (float)(Win - 1) / 2.0f; a.half_h = (float)(Hin - 1) / 2.0f;
    a.dwo = make_div((float)(W_orig - 1), div_mode); a.dho = make_div((float)(H_orig - 1), div_mode);
    a.rw = (float)((double)Win / (double)W_orig); a.rh = (float)((double)Hin / (double)H_orig);
    a.drw = make_div(a.rw, div_mode); a.drh = make_div(a.rh, div_mode);
    a.diff = (Hin != H_orig) || (Win != W_orig);
    return a;
}

// PixPro.py:140-143 + :168-175: centre of grid cell (x, y) of a crop descriptor, in pixels of the
// original frame:  ((i + 0.5) * ((c2 - c0) / G) + c0) * (size - 1)
__device__ __forceinline__ void grid_centre(const float* c, int x, int y, const ScalarDiv& dG, float wo, float ho, float& vx,
                                            float& vy) {
    const float bw = dG(sub(c[2], c[0])), bh = dG(sub(c[3], c[1]));
    const float fx = add((float)x, 0.5f), fy = add((float)y, 0.5f);
    vx = mul(add(mul(fx, bw), c[0]), wo);
    vy = mul(add(mul(fy, bh), c[1]), ho);
}

}  // namespace pp
