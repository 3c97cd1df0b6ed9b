// sta_sampler.cu — one launch for the elementwise part of a PLMS / DDIM sampler step, forward and backward.
//
// Reference ldm/models/diffusion/plms.py:296-358 (p_sample_plms): classifier-free guidance of the two UNet output rows
// (:304-308), the pseudo linear multistep (Adams-Bashforth) combination of the current and up to three previous noise
// estimates (:341-354) and the DDIM-style x_{t-1} update with eta = 0 (:321-338) — about a dozen elementwise torch launches
// per step and twenty more in its autograd backward, 150 times per image.  All of it is linear in (eps, x, old_eps):
//     e_t    = (1 - s) eps_u + s eps_c
//     e'     = w_e e_t + sum_k w_old[k] old_k
//     x_prev = a_x x + a_e e'                 a_x = sqrt(a_prev / a_t),  a_e = sqrt(1 - a_prev) - sqrt(a_prev (1 - a_t) / a_t)
//     pred_x0 = (x - sqrt(1 - a_t) e') / sqrt(a_t)        (returned for API compatibility, no gradient)
// so one fp32 kernel computes it and one computes the gradients.  HBM/latency-bound: 16 k elements at the 64x64 latent.
#include "../../include/sta_b200.h"
#include "sta_host.h"

namespace sta {

struct PlmsParams {
  const float4* eps_u;
  const float4* eps_c;
  const float4* x;
  const float4* old[3];
  float4* e_t;
  float4* x_prev;
  float4* pred_x0;
  long long n4;
  float s, w_e, w_old[3], a_x, a_e, p_x, p_e;
};

__device__ __forceinline__ float4 f4_axpby(float a, float4 x, float b, float4 y) {
  return make_float4(fmaf(a, x.x, b * y.x), fmaf(a, x.y, b * y.y), fmaf(a, x.z, b * y.z), fmaf(a, x.w, b * y.w));
}
__device__ __forceinline__ float4 f4_fma(float a, float4 x, float4 y) {
  return make_float4(fmaf(a, x.x, y.x), fmaf(a, x.y, y.y), fmaf(a, x.z, y.z), fmaf(a, x.w, y.w));
}
__device__ __forceinline__ float4 f4_scale(float a, float4 x) { return make_float4(a * x.x, a * x.y, a * x.z, a * x.w); }

__global__ void plms_step_fwd_kernel(const PlmsParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n4) return;
  const float4 eu = p.eps_u[i], ec = p.eps_c[i], x = p.x[i];
  const float4 et = f4_axpby(1.f - p.s, eu, p.s, ec);
  float4 ep = f4_scale(p.w_e, et);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (p.old[k]) ep = f4_fma(p.w_old[k], p.old[k][i], ep);
  p.e_t[i] = et;
  p.x_prev[i] = f4_axpby(p.a_x, x, p.a_e, ep);
  if (p.pred_x0) p.pred_x0[i] = f4_axpby(p.p_x, x, p.p_e, ep);
}

struct PlmsBwdParams {
  const float4* g_x_prev;  // may be null (= 0)
  const float4* g_e_t;     // may be null (= 0)
  float4* g_eps_u;
  float4* g_eps_c;
  float4* g_x;
  float4* g_old[3];  // null where the forward had no old_k
  long long n4;
  float s, w_e, w_old[3], a_x, a_e;
};

__global__ void plms_step_bwd_kernel(const PlmsBwdParams p) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n4) return;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 gx = p.g_x_prev ? p.g_x_prev[i] : z;
  const float4 ge = p.g_e_t ? p.g_e_t[i] : z;
  const float4 gp = f4_scale(p.a_e, gx);          // d e'
  const float4 gt = f4_fma(p.w_e, gp, ge);        // d e_t (both uses)
  p.g_eps_u[i] = f4_scale(1.f - p.s, gt);
  p.g_eps_c[i] = f4_scale(p.s, gt);
  p.g_x[i] = f4_scale(p.a_x, gx);
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (p.g_old[k]) p.g_old[k][i] = f4_scale(p.w_old[k], gp);
}

static bool aligned16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace sta

extern "C" int sta_plms_step_fwd(const sta_plms_step_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->eps || !a->x || !a->e_t || !a->x_prev) return fail(STA_ERR_BAD_ARG, "sta_plms_step_fwd: null pointer");
  if (a->prompts < 1 || a->elems < 4 || (a->elems % 4)) return fail(STA_ERR_UNSUPPORTED, "sta_plms_step_fwd: elems must be a positive multiple of 4");
  const void* ptrs[] = {a->eps, a->x, a->e_t, a->x_prev, a->pred_x0, a->old[0], a->old[1], a->old[2]};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(STA_ERR_UNSUPPORTED, "sta_plms_step_fwd: tensors must be 16-byte aligned");
  PlmsParams p;
  const long long n = (long long)a->prompts * a->elems;
  p.eps_u = reinterpret_cast<const float4*>(a->eps);
  p.eps_c = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a->eps) + n);
  p.x = reinterpret_cast<const float4*>(a->x);
  for (int k = 0; k < 3; ++k) { p.old[k] = reinterpret_cast<const float4*>(a->old[k]); p.w_old[k] = a->w_old[k]; }
  p.e_t = reinterpret_cast<float4*>(a->e_t);
  p.x_prev = reinterpret_cast<float4*>(a->x_prev);
  p.pred_x0 = reinterpret_cast<float4*>(a->pred_x0);
  p.n4 = n / 4;
  p.s = a->guidance; p.w_e = a->w_e; p.a_x = a->a_x; p.a_e = a->a_e; p.p_x = a->p_x; p.p_e = a->p_e;
  plms_step_fwd_kernel<<<(unsigned)((p.n4 + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}

extern "C" int sta_plms_step_bwd(const sta_plms_step_bwd_args* a, void* stream) {
  using namespace sta;
  if (!a || !a->g_eps || !a->g_x) return fail(STA_ERR_BAD_ARG, "sta_plms_step_bwd: null pointer");
  if (a->prompts < 1 || a->elems < 4 || (a->elems % 4)) return fail(STA_ERR_UNSUPPORTED, "sta_plms_step_bwd: elems must be a positive multiple of 4");
  const void* ptrs[] = {a->g_x_prev, a->g_e_t, a->g_eps, a->g_x, a->g_old[0], a->g_old[1], a->g_old[2]};
  for (const void* q : ptrs)
    if (q && !aligned16(q)) return fail(STA_ERR_UNSUPPORTED, "sta_plms_step_bwd: tensors must be 16-byte aligned");
  PlmsBwdParams p;
  const long long n = (long long)a->prompts * a->elems;
  p.g_x_prev = reinterpret_cast<const float4*>(a->g_x_prev);
  p.g_e_t = reinterpret_cast<const float4*>(a->g_e_t);
  p.g_eps_u = reinterpret_cast<float4*>(a->g_eps);
  p.g_eps_c = reinterpret_cast<float4*>(reinterpret_cast<float*>(a->g_eps) + n);
  p.g_x = reinterpret_cast<float4*>(a->g_x);
  for (int k = 0; k < 3; ++k) { p.g_old[k] = reinterpret_cast<float4*>(a->g_old[k]); p.w_old[k] = a->w_old[k]; }
  p.n4 = n / 4;
  p.s = a->guidance; p.w_e = a->w_e; p.a_x = a->a_x; p.a_e = a->a_e;
  plms_step_bwd_kernel<<<(unsigned)((p.n4 + 127) / 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  STA_CUDA_CHECK(cudaGetLastError());
  return STA_OK;
}
