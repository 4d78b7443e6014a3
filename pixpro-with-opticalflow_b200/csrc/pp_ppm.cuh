// pp_ppm.cuh — shared pieces of the PPM kernels (generic path pp_ppm.cu, small-grid path
// pp_ppm_small.cu).
#pragma once
#include "pp_common.cuh"

namespace pp {

constexpr float kNormEps = 1e-12f;  // F.normalize eps

// relu^γ and its derivative (contrast/models/PixPro.py:355-358)
struct Act {
    float gamma, cv;
    int mode;  // 0: γ==1, 1: γ==2, 2: general
    __device__ __forceinline__ float f(float s) const {
        float a = fmaxf(s, cv);
        if (gamma < 1.0f) a += 1e-6f;
        return mode == 0 ? a : (mode == 1 ? a * a : powf(a, gamma));
    }
    __device__ __forceinline__ float df(float s) const {  // d f / d s; torch clamp passes grad where s >= min
        if (s < cv) return 0.0f;
        float a = s;
        if (gamma < 1.0f) a += 1e-6f;
        return mode == 0 ? 1.0f : (mode == 1 ? 2.0f * a : gamma * powf(a, gamma - 1.0f));
    }
};
static inline Act make_act(double gamma, double cv) {
    Act a;
    a.gamma = (float)gamma;
    a.cv = (float)cv;
    a.mode = gamma == 1.0 ? 0 : (gamma == 2.0 ? 1 : 2);
    return a;
}

// small-grid path (pp_ppm_small.cu): one thread block per sample, everything in shared memory
bool ppm_small_supported(int C, int P);
int ppm_fwd_small(const float* feat, const float* val, int64_t B, int C, int P, Act act, int final_norm, float* out,
                  float* nx, float* nv, float* ny, float* S, cudaStream_t st);
int ppm_bwd_small(const float* feat, const float* val, const float* out, const float* g, int64_t B, int C, int P, Act act,
                  int final_norm, const float* nx, const float* nv, const float* ny, const float* S, float* d_feat,
                  float* d_val, cudaStream_t st);

}  // namespace pp
