// Fused multi-tensor Adam step for the optimizer that sits right behind the fusion block (SURVEY.md 8f rank 1;
// solver.py:30,91-94 uses torch.optim.Adam over model.parameters()).  One launch updates up to 32 parameter tensors:
// p, m, v are read and written once (7 x 4 bytes per element, the compulsory traffic of Adam) and the kernel ALSO writes
// the bf16 copy of the new weights that the tcgen05 GEMMs of the next step consume, so the per-step re-cast of 99 M
// parameters (a second read of every weight, 18 pack launches) disappears.
#include <cuda_bf16.h>

#include "common.h"

namespace vqa {
namespace {

constexpr int kMaxT = 32;
constexpr int kChunk = 4096;          // elements per block iteration (256 threads x 4 float4)

struct AdamTensors {
  float* p[kMaxT];
  const float* g[kMaxT];
  float* m[kMaxT];
  float* v[kMaxT];
  __nv_bfloat16* pb[kMaxT];           // optional bf16 copy of the updated parameter (nullptr: none)
  long long numel[kMaxT];
  long long start[kMaxT + 1];         // prefix sum of chunks
  int vec[kMaxT];                     // 1: every pointer 16-byte aligned (bf16: 8-byte) -> 128-bit path
  int n;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float step_size, float omb1,
                                         float beta2, float omb2, float inv_bc2_sqrt, float eps) {
  // torch/aten fused_adam_utils.cuh (non-amsgrad, weight_decay == 0, maximize == false).  ATen evaluates 1 - beta in
  // double (0.001, not the float 1.0000467e-3): omb1 / omb2 are those doubles rounded once, on the host.
  m = m + (g - m) * omb1;                          // lerp(exp_avg, grad, 1 - beta1)
  v = beta2 * v + omb2 * g * g;
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = p - step_size * m / denom;
}

// step_dev != nullptr: the 1-based step count lives on the device (a CUDA graph that contains this launch increments it
// itself before the update); the bias corrections are then formed here, in double like the host path.
__global__ void __launch_bounds__(256) adam_kernel(const AdamTensors T, float step_size, float omb1, float beta2,
                                                   float omb2, float inv_bc2_sqrt, float eps,
                                                   const long long* __restrict__ step_dev, double lr, double beta1_d,
                                                   double beta2_d) {
  if (step_dev != nullptr) {
    __shared__ float s_corr[2];
    if (threadIdx.x == 0) {
      const double st = (double)*step_dev;
      s_corr[0] = (float)(lr / (1.0 - pow(beta1_d, st)));
      s_corr[1] = (float)(1.0 / sqrt(1.0 - pow(beta2_d, st)));
    }
    __syncthreads();
    step_size = s_corr[0];
    inv_bc2_sqrt = s_corr[1];
  }
  const long long total = T.start[T.n];
  for (long long c = blockIdx.x; c < total; c += gridDim.x) {
    int i = 0;
    while (i + 1 < T.n && c >= T.start[i + 1]) ++i;
    const long long base = (c - T.start[i]) * kChunk;
    const long long n = T.numel[i];
    float* p = T.p[i];
    const float* g = T.g[i];
    float* m = T.m[i];
    float* v = T.v[i];
    __nv_bfloat16* pb = T.pb[i];
    if (T.vec[i] && base + kChunk <= n) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const long long e = base + (long long)(r * 256 + threadIdx.x) * 4;
        float4 P = *reinterpret_cast<const float4*>(p + e);
        const float4 G = __ldcs(reinterpret_cast<const float4*>(g + e));
        float4 M = *reinterpret_cast<const float4*>(m + e);
        float4 V = *reinterpret_cast<const float4*>(v + e);
        adam_one(P.x, G.x, M.x, V.x, step_size, omb1, beta2, omb2, inv_bc2_sqrt, eps);
        adam_one(P.y, G.y, M.y, V.y, step_size, omb1, beta2, omb2, inv_bc2_sqrt, eps);
        adam_one(P.z, G.z, M.z, V.z, step_size, omb1, beta2, omb2, inv_bc2_sqrt, eps);
        adam_one(P.w, G.w, M.w, V.w, step_size, omb1, beta2, omb2, inv_bc2_sqrt, eps);
        *reinterpret_cast<float4*>(p + e) = P;
        *reinterpret_cast<float4*>(m + e) = M;
        *reinterpret_cast<float4*>(v + e) = V;
        if (pb != nullptr) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(P.x, P.y), hi = __floats2bfloat162_rn(P.z, P.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&lo);
          u.y = *reinterpret_cast<uint32_t*>(&hi);
          *reinterpret_cast<uint2*>(pb + e) = u;
        }
      }
    } else {
      for (long long e = base + threadIdx.x; e < base + kChunk && e < n; e += 256) {
        float P = p[e], M = m[e], V = v[e];
        adam_one(P, g[e], M, V, step_size, omb1, beta2, omb2, inv_bc2_sqrt, eps);
        p[e] = P;
        m[e] = M;
        v[e] = V;
        if (pb != nullptr) pb[e] = __float2bfloat16_rn(P);
      }
    }
  }
}

}  // namespace
}  // namespace vqa

using namespace vqa;

// See include/vqa_b200.h.  Pointer tables are HOST arrays of device pointers (copied into the kernel's parameter block).
static int adam_impl(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                     void* const* exp_avg_sq, void* const* params_bf16, const int64_t* numel, double lr, double beta1,
                     double beta2, double eps, int64_t step, const int64_t* step_dev, void* stream) {
  if (n_tensors <= 0 || !params || !grads || !exp_avg || !exp_avg_sq || !numel || (step <= 0 && !step_dev))
    return set_error(VQA_B200_EINVAL, "adam_step: bad arguments");
  // hyper-parameters arrive as doubles (Python floats) and every derived constant is formed in double, as ATen does
  const double bc1 = 1.0 - pow(beta1, (double)(step > 0 ? step : 1));
  const double bc2 = 1.0 - pow(beta2, (double)(step > 0 ? step : 1));
  const float step_size = (float)(lr / bc1);
  const float inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int first = 0; first < n_tensors; first += kMaxT) {
    AdamTensors T;
    T.n = n_tensors - first < kMaxT ? n_tensors - first : kMaxT;
    long long chunks = 0;
    for (int i = 0; i < T.n; ++i) {
      const int k = first + i;
      if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || numel[k] < 0)
        return set_error(VQA_B200_EINVAL, "adam_step: null pointer in tensor %d", k);
      T.p[i] = (float*)params[k];
      T.g[i] = (const float*)grads[k];
      T.m[i] = (float*)exp_avg[k];
      T.v[i] = (float*)exp_avg_sq[k];
      T.pb[i] = params_bf16 ? (__nv_bfloat16*)params_bf16[k] : nullptr;
      T.numel[i] = numel[k];
      T.vec[i] = aligned16(T.p[i]) && aligned16(T.g[i]) && aligned16(T.m[i]) && aligned16(T.v[i]) &&
                 ((reinterpret_cast<uintptr_t>(T.pb[i]) & 7) == 0);
      T.start[i] = chunks;
      chunks += (numel[k] + kChunk - 1) / kChunk;
    }
    T.start[T.n] = chunks;
    if (chunks == 0) continue;
    const long long cap = (long long)sm_count() * 8;
    const int grid = (int)(chunks < cap ? chunks : cap);
    adam_kernel<<<grid, 256, 0, st>>>(T, step_size, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2),
                                      inv_bc2_sqrt, (float)eps, reinterpret_cast<const long long*>(step_dev), lr, beta1,
                                      beta2);
    VQA_LAUNCH_CHECK("adam_step");
  }
  return 0;
}

extern "C" int vqa_b200_adam_step(int n_tensors, void* const* params, const void* const* grads, void* const* exp_avg,
                                  void* const* exp_avg_sq, void* const* params_bf16, const int64_t* numel, double lr,
                                  double beta1, double beta2, double eps, int64_t step, void* stream) {
  return adam_impl(n_tensors, params, grads, exp_avg, exp_avg_sq, params_bf16, numel, lr, beta1, beta2, eps, step,
                   nullptr, stream);
}

extern "C" int vqa_b200_adam_step_dev(int n_tensors, void* const* params, const void* const* grads,
                                      void* const* exp_avg, void* const* exp_avg_sq, void* const* params_bf16,
                                      const int64_t* numel, double lr, double beta1, double beta2, double eps,
                                      const int64_t* step_dev, void* stream) {
  if (!step_dev) return set_error(VQA_B200_EINVAL, "adam_step_dev: null step counter");
  return adam_impl(n_tensors, params, grads, exp_avg, exp_avg_sq, params_bf16, numel, lr, beta1, beta2, eps, 0,
                   step_dev, stream);
}
