// HBM-bound kernels of the fusion / co-attention path: packing, attention logits, softmax +
// multi-glimpse pooling (forward / backward), the MFB elementwise backward and small helpers.
// All of them are single-pass over their large operand with 128-bit accesses.
#include "common.h"
#include "ptx.cuh"

namespace vqa {

// =====================================================================================
// packing
// =====================================================================================
__global__ void pack_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long d0,
                                 long long d1, long long d2, long long s0, long long s1, long long s2, long long t0,
                                 long long t1, int vec) {
  const long long total = d0 * d1 * d2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {   // s2 == 1, d2 % 8 == 0, all row starts 16-byte aligned
    const long long nv = total / 8;
    const long long d2v = d2 / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
      const long long c = i % d2v;
      const long long r = i / d2v;
      const long long b = r % d1;
      const long long a = r / d1;
      const float4* p = reinterpret_cast<const float4*>(src + a * s0 + b * s1 + c * 8);
      const float4 x = __ldg(p), y = __ldg(p + 1);
      uint4 u;
      u.x = pack_bf16(x.x, x.y); u.y = pack_bf16(x.z, x.w);
      u.z = pack_bf16(y.x, y.y); u.w = pack_bf16(y.z, y.w);
      *reinterpret_cast<uint4*>(dst + a * t0 + b * t1 + c * 8) = u;
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
      const long long c = i % d2;
      const long long r = i / d2;
      const long long b = r % d1;
      const long long a = r / d1;
      dst[a * t0 + b * t1 + c] = __float2bfloat16_rn(src[a * s0 + b * s1 + c * s2]);
    }
  }
}

// bf16 hi/lo split written three times along the contraction axis (see vqa_b200.h); batched, strided
__global__ void split3_kernel(const float* __restrict__ src, long long lds, long long sbs,
                              __nv_bfloat16* __restrict__ dst, long long ldd, long long dbs, long long batch,
                              long long R, long long C, int role, int concat_rows) {
  const long long total = batch * R * C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const long long c = i % C;
    const long long r = (i / C) % R;
    const long long bz = i / (C * R);
    const float x = src[bz * sbs + r * lds + c];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const __nv_bfloat16 s0 = hi;
    const __nv_bfloat16 s1 = role == 0 ? hi : lo;
    const __nv_bfloat16 s2 = role == 0 ? lo : hi;
    __nv_bfloat16* d = dst + bz * dbs;
    if (concat_rows) {
      d[(0 * R + r) * ldd + c] = s0;
      d[(1 * R + r) * ldd + c] = s1;
      d[(2 * R + r) * ldd + c] = s2;
    } else {
      d[r * ldd + 0 * C + c] = s0;
      d[r * ldd + 1 * C + c] = s1;
      d[r * ldd + 2 * C + c] = s2;
    }
  }
}

__global__ void dropout_mask_kernel(float* __restrict__ mask, int M, int N, uint32_t seed,
                                    const uint32_t* __restrict__ seed_dev, uint32_t thresh16, float scale) {
  seed = effective_seed(seed, seed_dev);
  const long long total = (long long)M * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t c = (uint32_t)(i % N);
    const uint32_t r = (uint32_t)(i / N);
    mask[i] = (thresh16 == 0 || dropout_keep(seed, r, c, thresh16)) ? scale : 0.f;
  }
}

// cp.async of one 8- or 16-byte piece (thread-private staging: the issuing thread is the only reader)
template <int BYTES>
__device__ __forceinline__ void cp_async_piece(void* smem_dst, const void* gsrc) {
  static_assert(BYTES == 8 || BYTES == 16, "cp.async piece");
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit_group() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// 4 elements out of a staged piece (bf16: 8 bytes, fp32: 16 bytes)
template <bool BF>
__device__ __forceinline__ void ld4_smem(const void* p, float* out) {
  if (BF) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    out[0] = bf16_lo(u.x); out[1] = bf16_hi(u.x); out[2] = bf16_lo(u.y); out[3] = bf16_hi(u.y);
  } else {
    const float4 a = *reinterpret_cast<const float4*>(p);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  }
}

// 4 consecutive elements (bf16: one 64-bit access, fp32: one 128-bit access)
template <bool BF>
__device__ __forceinline__ void ld4(const void* base, long long idx, float* out) {
  if (BF) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    out[0] = bf16_lo(u.x); out[1] = bf16_hi(u.x); out[2] = bf16_lo(u.y); out[3] = bf16_hi(u.y);
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
  }
}
template <bool BF>
__device__ __forceinline__ void st4(void* base, long long idx, const float* v) {
  if (BF) {
    uint2 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  }
}


// =====================================================================================
// attention logits: logits[m, g] = sum_j H[m, j] W2[g, j] + b2[g]        (warp per row)
// =====================================================================================
// bf16 rows with J <= 256*CH: the lane owns the columns j = lane*8 + 256*c (c < CH); its slice of W2 lives in
// registers, so the row loop touches no shared memory; two rows are in flight per warp.
template <int CH>
__global__ void __launch_bounds__(256) attn_logits_fwd_regs_kernel(const __nv_bfloat16* __restrict__ H, long long ldh,
                                                                   const float* __restrict__ W2,
                                                                   const float* __restrict__ b2,
                                                                   float* __restrict__ logits, int M, int J, int G) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  float wr[2][CH][8];
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int j = lane * 8 + 256 * c + e;
        wr[g][c][e] = (g < G && j < J) ? __ldg(W2 + g * J + j) : 0.f;
      }
  const float bias0 = b2[0], bias1 = G > 1 ? b2[1] : 0.f;
  const int stride = gridDim.x * warps * 2;
  for (int m = (blockIdx.x * warps + (threadIdx.x >> 5)) * 2; m < M; m += stride) {
    uint4 u[2][CH];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int mm = min(m + r, M - 1);
#pragma unroll
      for (int c = 0; c < CH; ++c)
        u[r][c] = (lane * 8 + 256 * c < J) ? __ldg(reinterpret_cast<const uint4*>(H + (long long)mm * ldh + lane * 8 + 256 * c))
                                           : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const uint32_t w[4] = {u[r][c].x, u[r][c].y, u[r][c].z, u[r][c].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x0 = bf16_lo(w[q]), x1 = bf16_hi(w[q]);
          a0 += x0 * wr[0][c][2 * q] + x1 * wr[0][c][2 * q + 1];
          a1 += x0 * wr[1][c][2 * q] + x1 * wr[1][c][2 * q + 1];
        }
      }
      a0 = warp_sum(a0);
      if (G > 1) a1 = warp_sum(a1);
      if (lane == 0 && m + r < M) {
        logits[(long long)(m + r) * G] = a0 + bias0;
        if (G > 1) logits[(long long)(m + r) * G + 1] = a1 + bias1;
      }
    }
  }
}

template <bool BF16>
__global__ void __launch_bounds__(256) attn_logits_fwd_kernel(const void* __restrict__ Hv, long long ldh,
                                                              const float* __restrict__ W2,
                                                              const float* __restrict__ b2,
                                                              float* __restrict__ logits, int M, int J, int G) {
  extern __shared__ float w2s[];   // [G][J]
  for (int i = threadIdx.x; i < G * J; i += blockDim.x) w2s[i] = W2[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int m = blockIdx.x * warps + (threadIdx.x >> 5); m < M; m += gridDim.x * warps) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (BF16) {
      const __nv_bfloat16* h = reinterpret_cast<const __nv_bfloat16*>(Hv) + (long long)m * ldh;
      for (int j = lane * 8; j < J; j += 256) {       // J % 8 == 0, rows 16-byte aligned
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(h + j));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float x0 = bf16_lo(w[q]), x1 = bf16_hi(w[q]);
#pragma unroll
          for (int g = 0; g < 4; ++g)
            if (g < G) acc[g] += x0 * w2s[g * J + j + 2 * q] + x1 * w2s[g * J + j + 2 * q + 1];
        }
      }
    } else {
      const float* h = reinterpret_cast<const float*>(Hv) + (long long)m * ldh;
      for (int j = lane; j < J; j += 32) {
        const float x = __ldg(h + j);
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g < G) acc[g] += x * w2s[g * J + j];
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      if (g < G) {
        const float sres = warp_sum(acc[g]);
        if (lane == 0) logits[(long long)m * G + g] = sres + b2[g];
      }
    }
  }
}

// V consecutive elements as 128-bit (or, for 4 bf16, 64-bit) accesses
template <bool BF, int V>
__device__ __forceinline__ void ldv(const void* base, long long idx, float* out) {
  if constexpr (BF && V == 8) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { out[2 * q] = bf16_lo(w[q]); out[2 * q + 1] = bf16_hi(w[q]); }
  } else {
#pragma unroll
    for (int v = 0; v < V; v += 4) ld4<BF>(base, idx + v, out + v);
  }
}
template <bool BF, int V>
__device__ __forceinline__ void stv(void* base, long long idx, const float* v) {
  if constexpr (BF && V == 8) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
#pragma unroll
    for (int q = 0; q < V; q += 4) st4<BF>(base, idx + q, v + q);
  }
}

// backward: a thread owns V adjacent columns (one 128-bit access: V = 8 for bf16 H and dH, else 4) of H / dH; the 256
// threads of a block form RL = 256 / (J / V) row lanes that walk a strip of rows; per-thread dW2 / dbias partials are
// reduced across the row lanes in shared memory and leave the block as atomics.  GN = compiled number of glimpses (2 for
// the co-attention / question attention, 4 = general): with 8-byte accesses, 4-glimpse register arrays and two rows in
// flight the kernel moved its 102 MB at 1.5 TB/s (latency-bound: 16 KB in flight per SM).
template <bool HBF16, bool DBF16, int V, int GN>
__global__ void __launch_bounds__(256) attn_logits_bwd_kernel(const void* __restrict__ Hv, long long ldh,
                                                              const float* __restrict__ W2,
                                                              const float* __restrict__ dlogits, void* __restrict__ dHv,
                                                              long long lddh, const float* __restrict__ out_scale,
                                                              int rows_per_group, int relu_mask,
                                                              float* __restrict__ dW2, float* __restrict__ db2,
                                                              float* __restrict__ dbias_h, int M, int J, int G,
                                                              int rows_per_block) {
  extern __shared__ float red[];                          // [RL][GN + 1][JT*V]: dW2[g < GN], dbias
  const int JT = (J + V - 1) / V;                         // threads along J (J % V == 0 checked by the launcher)
  const int RL = 256 / JT;                                // row lanes
  const int jt = threadIdx.x % JT, rl = threadIdx.x / JT;
  const int j0 = jt * V;
  const int m0 = blockIdx.x * rows_per_block;
  const int m1 = min(M, m0 + rows_per_block);
  float w2[GN][V], dw[GN][V], dbh[V];
  float db2_acc[GN];
#pragma unroll
  for (int g = 0; g < GN; ++g) {
    db2_acc[g] = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) { w2[g][v] = (g < G) ? W2[g * J + j0 + v] : 0.f; dw[g][v] = 0.f; }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) dbh[v] = 0.f;
  if (rl < RL) {
    constexpr int U = 4;                                  // rows in flight per thread
    for (int mb = m0 + rl; mb < m1; mb += U * RL) {
      float h[U][V], dl[U][GN], sc[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int m = min(mb + u * RL, m1 - 1);           // clamped rows are computed and discarded
        ldv<HBF16, V>(Hv, (long long)m * ldh + j0, h[u]);
#pragma unroll
        for (int g = 0; g < GN; ++g) dl[u][g] = (g < G) ? __ldg(dlogits + (long long)m * G + g) : 0.f;
        sc[u] = out_scale ? __ldg(out_scale + m / rows_per_group) : 1.f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int m = mb + u * RL;
        if (m >= m1) break;
        float d[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          float acc = 0.f;
#pragma unroll
          for (int g = 0; g < GN; ++g) { acc += dl[u][g] * w2[g][v]; dw[g][v] += dl[u][g] * h[u][v]; }
          if (relu_mask && !(h[u][v] > 0.f)) acc = 0.f;
          dbh[v] += acc;
          d[v] = acc * sc[u];
        }
        if (jt == 0) {
#pragma unroll
          for (int g = 0; g < GN; ++g) db2_acc[g] += dl[u][g];
        }
        stv<DBF16, V>(dHv, (long long)m * lddh + j0, d);
      }
    }
  }
  // ---- reduce over the row lanes
  const int W = JT * V;
  if (rl < RL) {
#pragma unroll
    for (int g = 0; g < GN; ++g)
#pragma unroll
      for (int v = 0; v < V; ++v) red[(rl * (GN + 1) + g) * W + j0 + v] = dw[g][v];
#pragma unroll
    for (int v = 0; v < V; ++v) red[(rl * (GN + 1) + GN) * W + j0 + v] = dbh[v];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (GN + 1) * W; i += 256) {
    const int q = i / W, j = i % W;
    if (j >= J || (q < GN && q >= G) || (q == GN && dbias_h == nullptr)) continue;
    float sres = 0.f;
    for (int r = 0; r < RL; ++r) sres += red[(r * (GN + 1) + q) * W + j];
    if (q < GN) atomicAdd(dW2 + q * J + j, sres);
    else atomicAdd(dbias_h + j, sres);
  }
  if (db2 != nullptr && jt == 0 && rl < RL) {
#pragma unroll
    for (int g = 0; g < GN; ++g)
      if (g < G) atomicAdd(db2 + g, db2_acc[g]);
  }
}

// =====================================================================================
// softmax over L + multi-glimpse pooling.
//   Row-contiguous streaming: a CTA owns a slice of consecutive region rows of ONE sample -- a single contiguous
//   span of HBM (rows x D elements, ~200 KB) -- and its 256 threads cover a whole row (thread t owns the 16 bytes at
//   column t*V of every row), so every warp-level request is a full 512-byte run and consecutive requests walk
//   linearly through DRAM pages.  U independent 128-bit loads per thread are in flight.
// =====================================================================================
// Forward: one CTA (NW warps) per (sample, 32*V-column chunk).  Each warp streams 1/NW of the region rows of that
// chunk with U independent 128-bit loads in flight per lane and accumulates every glimpse in registers; the NW partial
// sums meet in shared memory.  NW is chosen at launch so that the whole grid (N * D / (32*V) CTAs: 2048 at N=256,
// D=2048, bf16) is ONE resident wave: at 62 registers per thread an SM holds 32/NW CTAs, and with 4 warps per CTA the
// 2048 CTAs were 1.73 waves -- the ragged second wave left a quarter of the machine idle at the end (ncu).
template <bool BF16, int G, int NW>
__global__ void __launch_bounds__(NW * 32) softmax_pool_fwd_kernel(const void* __restrict__ Xv,
                                                               const float* __restrict__ logits,
                                                               float* __restrict__ att, float* __restrict__ pooled,
                                                               int L, int D, int chunks, int degenerate) {
  constexpr int V = BF16 ? 8 : 4;
  constexpr int ES = BF16 ? 2 : 4;
  extern __shared__ float sm[];
  float* w = sm;                            // [G][L]
  float* part = sm + G * L;                 // [NW-1][G][32*V] partial sums of warps 1..NW-1
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  // --- the first batch of feature loads goes out BEFORE the softmax prologue: they do not depend on the weights, and
  // with the whole grid being one resident wave every CTA would otherwise leave DRAM idle for the ~2 us of its prologue
  constexpr int U = 7;
  const int d0 = chunk * 32 * V + lane * V;
  const bool act = d0 < D;
  const char* xb = reinterpret_cast<const char*>(Xv) + ((long long)n * L * D + (act ? d0 : 0)) * ES;
  const long long pitch = (long long)D * ES;
  const int l_end = (int)(((long long)(warp + 1) * L) / NW);
  int l = (int)(((long long)warp * L) / NW);
  uint4 buf[U];
  const bool first_full = act && (l + U <= l_end);
  if (first_full) {
#pragma unroll
    for (int u = 0; u < U; ++u) buf[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(l + u) * pitch));
  }
  // --- softmax over L: warp (g mod NW) computes glimpse g
  for (int g = warp; g < G; g += NW) {
    if (degenerate) {
      for (int l = lane; l < L; l += 32) w[g * L + l] = 1.f;
    } else {
      float mx = -INFINITY;
      for (int l = lane; l < L; l += 32) mx = fmaxf(mx, __ldg(logits + ((long long)n * L + l) * G + g));
      mx = warp_max(mx);
      float ssum = 0.f;
      for (int l = lane; l < L; l += 32) {
        const float e = __expf(__ldg(logits + ((long long)n * L + l) * G + g) - mx);
        w[g * L + l] = e;
        ssum += e;
      }
      ssum = warp_sum(ssum);
      const float r = 1.f / ssum;
      for (int l = lane; l < L; l += 32) w[g * L + l] *= r;
    }
    if (chunk == 0 && att != nullptr)
      for (int l = lane; l < L; l += 32) att[((long long)n * G + g) * L + l] = w[g * L + l];
  }
  __syncthreads();
  // --- pooling: warp q streams rows [q*L/NW, (q+1)*L/NW) of X[n, :, chunk]
  float acc[G][V];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int v = 0; v < V; ++v) acc[g][v] = 0.f;
  if (act) {
    bool have = first_full;                 // the first batch is already in registers (issued above)
    for (; l + U <= l_end; l += U) {
      if (!have) {
#pragma unroll
        for (int u = 0; u < U; ++u) buf[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)(l + u) * pitch));
      }
      have = false;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t uu[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float a = w[g * L + l + u];
          if (BF16) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              acc[g][(2 * q) % V] += a * bf16_lo(uu[q]);
              acc[g][(2 * q + 1) % V] += a * bf16_hi(uu[q]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[g][q % V] += a * __uint_as_float(uu[q]);
          }
        }
      }
    }
    for (; l < l_end; ++l) {
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(xb + (long long)l * pitch));
      const uint32_t uu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float a = w[g * L + l];
        if (BF16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[g][(2 * q) % V] += a * bf16_lo(uu[q]);
            acc[g][(2 * q + 1) % V] += a * bf16_hi(uu[q]);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[g][q % V] += a * __uint_as_float(uu[q]);
        }
      }
    }
  }
  if (warp > 0) {
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int v = 0; v < V; ++v) part[((warp - 1) * G + g) * 32 * V + lane * V + v] = acc[g][v];
  }
  __syncthreads();
  if (warp == 0 && act) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
#pragma unroll
      for (int v = 0; v < V; ++v)
#pragma unroll
        for (int q = 0; q + 1 < NW; ++q) acc[g][v] += part[(q * G + g) * 32 * V + lane * V + v];
      float* o = pooled + (long long)n * G * D + (long long)g * D + d0;
#pragma unroll
      for (int v = 0; v < V; v += 4) *reinterpret_cast<float4*>(o + v) = make_float4(acc[g][v], acc[g][v + 1], acc[g][v + 2], acc[g][v + 3]);
    }
  }
}

// Sums NV per-lane values over the 32 lanes of a warp with NV + log2(32/NV) - 1... shuffles instead of 5*NV: each
// exchange step also halves the number of values a lane carries (the upper half of the lanes keeps the upper half of the
// values), so that after log2(NV) steps every lane holds ONE value, index (lane >> log2(32/NV)) & (NV-1), which the
// remaining plain butterfly steps complete.  Returns that value's warp-wide sum.
template <int NV>
__device__ __forceinline__ float warp_multi_sum(float (&v)[NV], int lane) {
  static_assert(NV == 1 || NV == 2 || NV == 4 || NV == 8 || NV == 16, "NV must be a power of two <= 16");
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off >= 1; off >>= 1) v[0] += __shfl_xor_sync(0xFFFFFFFFu, v[0], off);
  return v[0];
}

// backward, pass A: datt[n,g,l] = sum_d dP[n,g,d] X[n,l,d]  (+ optional dX[n,l,:] = sum_g att[n,g,l] dP[n,g,:]).
// Same decomposition as the forward: one CTA (NW warps, one resident wave) per (sample, 512-byte column chunk), each
// warp streams 1/NW of the rows with U loads in flight; the lane keeps its 16-byte slice of dP in registers, reduces each row's partial
// dot with shuffles and adds it to datt (zero-initialised, [N, G, L] order inside the dlogits buffer; D / (32*V)
// chunks contribute to every element).  Pass B turns datt into dlogits in place.
template <bool BF16, int G, bool HAS_DX, int NW>
__global__ void __launch_bounds__(NW * 32) softmax_pool_bwd_kernel(const void* __restrict__ Xv,
                                                               const float* __restrict__ att,
                                                               const float* __restrict__ dpooled,
                                                               float* __restrict__ datt, float* __restrict__ dX,
                                                               const float* __restrict__ datt_extra,
                                                               unsigned int* __restrict__ done, int L, int D,
                                                               int chunks, int degenerate, int accumulate_dx) {
  constexpr int V = BF16 ? 8 : 4;
  constexpr int ES = BF16 ? 2 : 4;
  extern __shared__ float sm[];                                // pass B scratch of the sample's last CTA: [2][G][L] + [G]
  __shared__ int is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int d0 = min(chunk * 32 * V + lane * V, D - V);       // lanes past the end re-read the last slice ...
  const bool act = chunk * 32 * V + lane * V < D;             // ... and contribute nothing
  float dpv[G][V];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int v = 0; v < V; v += 4) {
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(dpooled + (long long)n * G * D + (long long)g * D + d0 + v));
      dpv[g][v] = act ? t4.x : 0.f; dpv[g][v + 1] = act ? t4.y : 0.f;
      dpv[g][v + 2] = act ? t4.z : 0.f; dpv[g][v + 3] = act ? t4.w : 0.f;
    }
  constexpr int U = 7;
  const char* xb = reinterpret_cast<const char*>(Xv) + ((long long)n * L * D + d0) * ES;
  const long long pitch = (long long)D * ES;
  const int l_end = (int)(((long long)(warp + 1) * L) / NW);
  for (int l0 = (int)(((long long)warp * L) / NW); l0 < l_end; l0 += U) {
    uint4 buf[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int l = min(l0 + u, l_end - 1);                   // clamped rows are computed and discarded
      buf[u] = __ldg(reinterpret_cast<const uint4*>(xb + (long long)l * pitch));
    }
    constexpr int NV = (U * G > 8) ? 16 : 8;          // per-lane partial dots of this batch of rows, padded
    float dots[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) dots[i] = 0.f;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float xv[V];
      const uint32_t uu[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
      if (BF16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { xv[(2 * q) % V] = bf16_lo(uu[q]); xv[(2 * q + 1) % V] = bf16_hi(uu[q]); }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) xv[q % V] = __uint_as_float(uu[q]);
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float dot = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) dot += xv[v] * dpv[g][v];
        dots[u * G + g] = dot;
      }
    }
    {
      // one transposing reduction for all U*G dots of the batch (16 shuffles instead of 70), then the lanes that end
      // up holding a dot add it to datt side by side
      const float tot = warp_multi_sum<NV>(dots, lane);
      const int idx = (lane / (32 / NV)) & (NV - 1);
      const int u = idx / G, g = idx % G, l = l0 + u;
      if ((lane & (32 / NV - 1)) == 0 && idx < U * G && l < l_end)
        atomicAdd(datt + ((long long)n * G + g) * L + l, tot);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int l = l0 + u;
      if (HAS_DX && act && l < l_end) {
        float* o = dX + ((long long)n * L + l) * D + d0;
        float aw[G];
#pragma unroll
        for (int g = 0; g < G; ++g) aw[g] = degenerate ? 1.f : __ldg(att + ((long long)n * G + g) * L + l);
#pragma unroll
        for (int v = 0; v < V; v += 4) {
          float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int g = 0; g < G; ++g) {
            res.x += aw[g] * dpv[g][v]; res.y += aw[g] * dpv[g][v + 1];
            res.z += aw[g] * dpv[g][v + 2]; res.w += aw[g] * dpv[g][v + 3];
          }
          if (accumulate_dx) {
            const float4 prev = *reinterpret_cast<const float4*>(o + v);
            res.x += prev.x; res.y += prev.y; res.z += prev.z; res.w += prev.w;
          }
          *reinterpret_cast<float4*>(o + v) = res;
        }
      }
    }
  }
  // --- pass B, folded in: the CTA that completes a sample's datt (the last of its `chunks` CTAs to get here) turns it
  // into dlogits in place: dlogits[n,l,g] = att * (datt - sum_l att * datt).  No second launch, no grid-wide wait.
  __threadfence();                                   // this CTA's atomics are visible before its ticket is taken
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(done + n, 1u) == (unsigned)(chunks - 1));
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  float* a_s = sm;                // [G][L]
  float* da_s = sm + G * L;       // [G][L]
  float* ssum = da_s + G * L;     // [G]
  float* bufn = datt + (long long)n * G * L;
  for (int i = tid; i < G * L; i += NW * 32) {
    a_s[i] = degenerate ? 1.f : att[(long long)n * G * L + i];
    float d = __ldcg(bufn + i);                      // written by other CTAs' atomics: read at L2
    if (datt_extra) d += datt_extra[(long long)n * G * L + i];
    da_s[i] = d;
  }
  __syncthreads();
  for (int g = warp; g < G; g += NW) {
    float sacc = 0.f;
    for (int l = lane; l < L; l += 32) sacc += a_s[g * L + l] * da_s[g * L + l];
    sacc = warp_sum(sacc);
    if (lane == 0) ssum[g] = sacc;
  }
  __syncthreads();
  for (int i = tid; i < G * L; i += NW * 32) {
    const int l = i / G, g = i % G;            // output order [L][G]
    bufn[i] = degenerate ? 0.f : a_s[g * L + l] * (da_s[g * L + l] - ssum[g]);
  }
}

// =====================================================================================
// MFB elementwise backward (see vqa_b200.h).  One thread owns 40 adjacent columns (= 8 pooled outputs): per row it
// moves 80/160 B of `keep` and of `dI` plus 16/32 B of y and g, all as 128-bit accesses.  grid (groups, row slices).
// =====================================================================================
template <bool BF>
__device__ __forceinline__ void ld8(const void* base, long long idx, float* out) {   // 8 consecutive elements
  if (BF) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + idx));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { out[2 * q] = bf16_lo(w[q]); out[2 * q + 1] = bf16_hi(w[q]); }
  } else {
    const float4 a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx));
    const float4 b = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + idx) + 1);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  }
}
template <bool BF>
__device__ __forceinline__ void st8(void* base, long long idx, const float* v) {
  if (BF) {
    uint4 u;
    u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = u;
  } else {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}

constexpr int MFB_BWD_PREFETCH = 4;      // rows staged ahead per thread in mfb_bwd_kernel

// One thread owns 20 adjacent columns (= 4 pooled outputs): 60 accumulator registers instead of 120, so four CTAs
// of 256 threads fit per SM (the 40-column version was register-bound at 8 warps/SM and latency-limited).
// CASCADE: the vector-block extras (several L2-norm segments, MHB's cascade operands); compiled out of the grid MFB's
// instance, whose inner loop they would slow down.
template <bool YG_BF16, bool KD_BF16, bool CASCADE>
__global__ void __launch_bounds__(256) mfb_bwd_kernel(const void* __restrict__ Gv, long long ldg,
                                                      const void* __restrict__ Yv, long long ldy,
                                                      const float* __restrict__ inv, const float* __restrict__ t,
                                                      const float* __restrict__ Q, long long ldq,
                                                      const void* __restrict__ keep, void* __restrict__ dIv,
                                                      float* __restrict__ dQ, float* __restrict__ dbias,
                                                      const float* __restrict__ extra,
                                                      const float* __restrict__ dprod_in, float* __restrict__ dExtra,
                                                      int rows_per_group, int rows_per_slice, int M, int N,
                                                      int seg_cols, uint32_t seed,
                                                      const uint32_t* __restrict__ seed_dev, uint32_t thresh16,
                                                      float scale) {
  seed = effective_seed(seed, seed_dev);
  const int grp = blockIdx.x;
  const int c0 = (blockIdx.z * 256 + threadIdx.x) * 20;      // blockIdx.z: 5120-column blocks (N = 10000: two MFB blocks)
  if (c0 >= N) return;
  const int o0 = c0 / 5;
  const int gi = CASCADE ? grp * (N / seg_cols) + c0 / seg_cols : grp;     // (group, L2-norm segment): seg_cols % 20 == 0
  const int g0 = grp * rows_per_group;
  const int m0 = g0 + blockIdx.y * rows_per_slice;
  const int m1 = min(min(M, g0 + rows_per_group), m0 + rows_per_slice);
  if (m0 >= m1) return;
  const float iv = inv[gi];
  const float coef = iv * iv * t[gi];
  // q = the effective multiplier Q * extra (MHB's cascade, mhb_coAtt.py:204-205); q0 / e keep the two factors apart for
  // the gradients dQ = S * extra and dExtra = S * Q with S = sum_m u * keep
  float q[20], dq[20], db[20];
#pragma unroll
  for (int i = 0; i < 20; i += 4) {
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(Q + (long long)grp * ldq + c0 + i));
    q[i] = q4.x; q[i + 1] = q4.y; q[i + 2] = q4.z; q[i + 3] = q4.w;
  }
  if (CASCADE && extra != nullptr) {
#pragma unroll
    for (int i = 0; i < 20; i += 4) {
      const float4 e4 = __ldg(reinterpret_cast<const float4*>(extra + (long long)grp * ldq + c0 + i));
      q[i] *= e4.x; q[i + 1] *= e4.y; q[i + 2] *= e4.z; q[i + 3] *= e4.w;
    }
  }
#pragma unroll
  for (int i = 0; i < 20; ++i) { dq[i] = 0.f; db[i] = 0.f; }
  // The rows of the thread's column strip (keep: 5 pieces, y and g: one piece each) are staged PF rows ahead through
  // THREAD-PRIVATE shared-memory slots with cp.async: the loads of four rows are in flight without holding registers
  // (the register-staged version sat at 25 % occupancy with half of its stall samples on the global loads), and since a
  // thread only ever reads what it copied itself no barrier is needed, just cp.async.wait_group.
  constexpr int PF = MFB_BWD_PREFETCH;
  constexpr int KP = KD_BF16 ? 8 : 16, YP = YG_BF16 ? 8 : 16, KE = KD_BF16 ? 2 : 4, YE = YG_BF16 ? 2 : 4;
  extern __shared__ __align__(16) unsigned char mfb_stage[];
  unsigned char* ks = mfb_stage;                                        // [PF][5][256] pieces of KP bytes
  unsigned char* ys = mfb_stage + (size_t)PF * 5 * 256 * KP;            // [PF][2][256] pieces of YP bytes
  const int tid = threadIdx.x;
  auto issue = [&](int m, int stage) {
    if (m < m1) {
      const char* kp = reinterpret_cast<const char*>(keep) + ((long long)m * N + c0) * KE;
#pragma unroll
      for (int i = 0; i < 5; ++i) cp_async_piece<KP>(ks + ((size_t)(stage * 5 + i) * 256 + tid) * KP, kp + i * KP);
      cp_async_piece<YP>(ys + ((size_t)(stage * 2 + 0) * 256 + tid) * YP,
                         reinterpret_cast<const char*>(Yv) + ((long long)m * ldy + o0) * YE);
      cp_async_piece<YP>(ys + ((size_t)(stage * 2 + 1) * 256 + tid) * YP,
                         reinterpret_cast<const char*>(Gv) + ((long long)m * ldg + o0) * YE);
    }
    cp_async_commit_group();             // committed even when empty: the group count stays one per row slot
  };
#pragma unroll
  for (int s_ = 0; s_ < PF; ++s_) issue(m0 + s_, s_);
  for (int m = m0; m < m1; ++m) {
    const int stage = (m - m0) % PF;
    cp_async_wait_group<PF - 1>();       // the oldest outstanding group (row m) has landed
    float y[4], g[4], kv[20];
    ld4_smem<YG_BF16>(ys + ((size_t)(stage * 2 + 0) * 256 + tid) * YP, y);
    ld4_smem<YG_BF16>(ys + ((size_t)(stage * 2 + 1) * 256 + tid) * YP, g);
#pragma unroll
    for (int i = 0; i < 5; ++i) ld4_smem<KD_BF16>(ks + ((size_t)(stage * 5 + i) * 256 + tid) * KP, kv + 4 * i);
    // the kernel is issue-bound (ncu: 477 warp instructions per row, 62 % of the issue slots, 46 % ALU pipe), so the
    // per-element work is kept to select / multiply / add / fma: dz and dz * scale are formed once per pooled output
    float dz[4], dzs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float ay = fabsf(y[j]);
      dz[j] = ay > 0.f ? __fdividef(g[j] - y[j] * coef, 2.f * ay) : 0.f;
      dzs[j] = dz[j] * scale;
    }
    float di[20];
#pragma unroll
    for (int i = 0; i < 20; i += 2) {
      bool k0 = true, k1 = true;
      if (thresh16) {
        const uint32_t rb = dropout_bits(seed, (uint32_t)m, (uint32_t)((c0 + i) >> 1));
        k0 = (rb & 0xFFFFu) >= thresh16;
        k1 = (rb >> 16) >= thresh16;
      }
      float d0 = dz[i / 5], d1 = dz[(i + 1) / 5];
      float u0 = dzs[i / 5], u1 = dzs[(i + 1) / 5];       // d * mask value of a kept element
      if (CASCADE && dprod_in != nullptr) {   // gradient arriving at the dropped-out product itself (next block of the cascade)
        const float2 dp = __ldg(reinterpret_cast<const float2*>(dprod_in + (long long)m * N + c0 + i));
        d0 += dp.x; d1 += dp.y;
        u0 = d0 * scale; u1 = d1 * scale;
      }
      u0 = k0 ? u0 : 0.f;
      u1 = k1 ? u1 : 0.f;
      di[i] = u0 * q[i];
      di[i + 1] = u1 * q[i + 1];
      dq[i] += d0 * kv[i];
      dq[i + 1] += d1 * kv[i + 1];
      db[i] += u0;
      db[i + 1] += u1;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) st4<KD_BF16>(dIv, (long long)m * N + c0 + 4 * i, di + 4 * i);
    issue(m + PF, stage);                // refill the slot that was just consumed
  }
  if (CASCADE && (extra != nullptr || dExtra != nullptr)) {
    // dq holds S = sum_m u * keep: dQ = S * extra, dExtra = S * Q  (only used with one slice per group)
#pragma unroll
    for (int i = 0; i < 20; i += 4) {
      const float4 q4 = __ldg(reinterpret_cast<const float4*>(Q + (long long)grp * ldq + c0 + i));
      if (dExtra != nullptr)
        *reinterpret_cast<float4*>(dExtra + (long long)grp * N + c0 + i) =
            make_float4(dq[i] * q4.x, dq[i + 1] * q4.y, dq[i + 2] * q4.z, dq[i + 3] * q4.w);
      if (extra != nullptr) {
        const float4 e4 = __ldg(reinterpret_cast<const float4*>(extra + (long long)grp * ldq + c0 + i));
        dq[i] *= e4.x; dq[i + 1] *= e4.y; dq[i + 2] *= e4.z; dq[i + 3] *= e4.w;
      }
    }
  }
  float* dqrow = dQ + (long long)grp * N + c0;
  if (gridDim.y == 1) {
#pragma unroll
    for (int i = 0; i < 20; i += 4) *reinterpret_cast<float4*>(dqrow + i) = make_float4(dq[i], dq[i + 1], dq[i + 2], dq[i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 20; i += 4)
      asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dqrow + i), "f"(dq[i]), "f"(dq[i + 1]),
                   "f"(dq[i + 2]), "f"(dq[i + 3])
                   : "memory");
  }
  if (dbias) {
#pragma unroll
    for (int i = 0; i < 20; i += 4)
      asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dbias + c0 + i), "f"(db[i] * q[i]),
                   "f"(db[i + 1] * q[i + 1]), "f"(db[i + 2] * q[i + 2]), "f"(db[i + 3] * q[i + 3])
                   : "memory");
  }
}

// g[m,o] = d[m,o] * inv[grp];  t[grp] += sum_o y[m,o] * g[m,o]            (warp per row)
template <bool YBF16>
__global__ void __launch_bounds__(256) norm_bwd_prep_kernel(const float* __restrict__ d, long long ldd,
                                                            const void* __restrict__ Yv, long long ldy,
                                                            const float* __restrict__ inv, float* __restrict__ g,
                                                            long long ldg, float* __restrict__ t,
                                                            int rows_per_group, int M, int No) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int m = blockIdx.x * warps + (threadIdx.x >> 5); m < M; m += gridDim.x * warps) {
    const int grp = m / rows_per_group;
    const float iv = inv[grp];
    float acc = 0.f;
    for (int o = lane; o < No; o += 32) {
      const float y = YBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(Yv)[(long long)m * ldy + o])
                            : reinterpret_cast<const float*>(Yv)[(long long)m * ldy + o];
      const float gv = d[(long long)m * ldd + o] * iv;
      g[(long long)m * ldg + o] = gv;
      acc += y * gv;
    }
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(t + grp, acc);
  }
}

__global__ void inv_norm_kernel(const float* __restrict__ ssq, float* __restrict__ inv, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[i] = 1.f / fmaxf(sqrtf(ssq[i]), 1e-12f);
}

template <bool YBF16>
__global__ void scale_rows_kernel(const void* __restrict__ Yv, long long ldy, const float* __restrict__ inv,
                                  int rows_per_group, float* __restrict__ out, long long ldo, int M, int No) {
  const long long total = (long long)M * No;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int o = (int)(i % No);
    const int m = (int)(i / No);
    const float y = YBF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(Yv)[(long long)m * ldy + o])
                          : reinterpret_cast<const float*>(Yv)[(long long)m * ldy + o];
    out[(long long)m * ldo + o] = y * inv[m / rows_per_group];
  }
}

// ---- small generic helpers (runtime dtype; uniform branch) ----
__device__ __forceinline__ float ld_any(const void* p, int bf16, long long i) {
  return bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void st_any(void* p, int bf16, long long i, float v) {
  if (bf16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}

// t[grp] += sum_o A[m,o] * B[m,o]                                        (warp per row)
__global__ void __launch_bounds__(256) group_dot_kernel(const void* __restrict__ A, int abf, long long lda,
                                                        const void* __restrict__ B, int bbf, long long ldb,
                                                        float* __restrict__ t, int rows_per_group, int M, int No) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (int m = blockIdx.x * warps + (threadIdx.x >> 5); m < M; m += gridDim.x * warps) {
    float acc = 0.f;
    for (int o = lane; o < No; o += 32) acc += ld_any(A, abf, (long long)m * lda + o) * ld_any(B, bbf, (long long)m * ldb + o);
    acc = warp_sum(acc);
    if (lane == 0) atomicAdd(t + m / rows_per_group, acc);
  }
}

// out[j] += sum_m X[m, j];   block = 64 columns x 4 row lanes, grid.y = row strips
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ X, int xbf, long long ldx,
                                                     float* __restrict__ out, int M, int J, int rows_per_block) {
  __shared__ float red[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + tx;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (j < J)
    for (int m = m0 + ty; m < m1; m += 4) acc += ld_any(X, xbf, (long long)m * ldx + j);
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && j < J) atomicAdd(out + j, red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx]);
}

// 128-bit form of colsum_kernel: a thread owns V adjacent columns (8 bf16 / 4 fp32), 32 threads cover a 512-byte row
// segment, 8 row lanes with four rows in flight each.  (The element-per-thread kernel above moved the 54 MB bf16 gate
// gradients of the recurrence at 1.2 TB/s: two bytes per thread and iteration.)
template <bool BF>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const void* __restrict__ X, long long ldx,
                                                         float* __restrict__ out, int M, int J, int rows_per_block) {
  constexpr int V = BF ? 8 : 4;
  __shared__ float red[8][32 * V];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j0 = (blockIdx.x * 32 + tx) * V;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  if (j0 < J) {
    for (int mb = m0 + ty; mb < m1; mb += 32) {
      float x[4][V];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int m = mb + 8 * u;
        if (m < m1) {
          ldv<BF, V>(X, (long long)m * ldx + j0, x[u]);
        } else {
#pragma unroll
          for (int v = 0; v < V; ++v) x[u][v] = 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += x[u][v];
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) red[ty][tx * V + v] = acc[v];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * V; i += 256) {
    const int j = blockIdx.x * 32 * V + i;
    if (j < J) {
      float sres = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) sres += red[r][i];
      atomicAdd(out + j, sres);
    }
  }
}

// out[m,j] = (H[m,j] > 0 ? D[m,j] : 0) * scale[m / rpg];  dbias[j] += unscaled masked D
__global__ void __launch_bounds__(256) relu_bwd_kernel(const void* __restrict__ D, int dbf, long long ldd,
                                                       const void* __restrict__ H, int hbf, long long ldh,
                                                       void* __restrict__ out, int obf, long long ldo,
                                                       const float* __restrict__ scale, int rows_per_group,
                                                       float* __restrict__ dbias, int M, int J, int rows_per_block) {
  __shared__ float red[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + tx;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (j < J) {
    for (int m = m0 + ty; m < m1; m += 4) {
      const float h = ld_any(H, hbf, (long long)m * ldh + j);
      float d = ld_any(D, dbf, (long long)m * ldd + j);
      if (!(h > 0.f)) d = 0.f;
      acc += d;
      st_any(out, obf, (long long)m * ldo + j, scale ? d * scale[m / rows_per_group] : d);
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (dbias && ty == 0 && j < J) atomicAdd(dbias + j, red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx]);
}

static int ew_grid(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)sm_count() * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace vqa

using namespace vqa;
#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int vqa_b200_pack_bf16(const float* src, void* dst, int64_t d0, int64_t d1, int64_t d2, int64_t s0,
                                  int64_t s1, int64_t s2, int64_t t0, int64_t t1, void* stream) {
  if (!src || !dst || d0 <= 0 || d1 <= 0 || d2 <= 0) return set_error(VQA_B200_EINVAL, "pack_bf16: bad arguments");
  if (t1 <= 0) t1 = d2;
  if (t0 <= 0) t0 = d1 * t1;
  const int vec = (s2 == 1) && (d2 % 8 == 0) && aligned16(src) && aligned16(dst) && (s0 % 4 == 0) && (s1 % 4 == 0) &&
                  (t0 % 8 == 0) && (t1 % 8 == 0);
  const long long work = vec ? d0 * d1 * d2 / 8 : d0 * d1 * d2;
  pack_bf16_kernel<<<ew_grid(work, 256), 256, 0, ST(stream)>>>(src, reinterpret_cast<__nv_bfloat16*>(dst), d0, d1, d2,
                                                              s0, s1, s2, t0, t1, vec);
  VQA_LAUNCH_CHECK("pack_bf16");
  return 0;
}

extern "C" int vqa_b200_split3_bf16(const float* src, int64_t lds, int64_t src_bstride, void* dst, int64_t ldd,
                                    int64_t dst_bstride, int64_t batch, int64_t R, int64_t C, int role,
                                    int concat_rows, void* stream) {
  if (!src || !dst || R <= 0 || C <= 0 || batch <= 0) return set_error(VQA_B200_EINVAL, "split3_bf16: bad arguments");
  split3_kernel<<<ew_grid(batch * R * C, 256), 256, 0, ST(stream)>>>(src, lds, src_bstride,
                                                                    reinterpret_cast<__nv_bfloat16*>(dst), ldd,
                                                                    dst_bstride, batch, R, C, role, concat_rows);
  VQA_LAUNCH_CHECK("split3_bf16");
  return 0;
}

static void drop_params(float p, uint32_t* thresh16, float* scale) {
  *thresh16 = (uint32_t)(p * 65536.0f + 0.5f);
  *scale = *thresh16 ? 65536.0f / (65536.0f - (float)*thresh16) : 1.0f;
}

extern "C" int vqa_b200_dropout_mask(float* mask, int M, int N, float drop_p, uint32_t seed, const uint32_t* seed_dev,
                                     void* stream) {
  if (!mask || M <= 0 || N <= 0) return set_error(VQA_B200_EINVAL, "dropout_mask: bad arguments");
  uint32_t th; float sc;
  drop_params(drop_p, &th, &sc);
  dropout_mask_kernel<<<ew_grid((long long)M * N, 256), 256, 0, ST(stream)>>>(mask, M, N, seed, seed_dev, th, sc);
  VQA_LAUNCH_CHECK("dropout_mask");
  return 0;
}

extern "C" int vqa_b200_attn_logits_fwd(const void* H, int h_dtype, int64_t ldh, const float* W2, const float* b2,
                                        float* logits, int M, int J, int G, void* stream) {
  if (!H || !W2 || !b2 || !logits || M <= 0 || J <= 0 || G <= 0 || G > 4)
    return set_error(VQA_B200_EINVAL, "attn_logits_fwd: bad arguments (G must be 1..4)");
  const size_t smem = (size_t)G * J * sizeof(float);
  int grid = (M + 7) / 8;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  if (h_dtype == VQA_B200_BF16) {
    if (J % 8 != 0 || !aligned16(H) || (ldh * 2) % 16 != 0)
      return set_error(VQA_B200_EALIGN, "attn_logits_fwd: bf16 H needs J %% 8 == 0 and 16-byte aligned rows");
    const __nv_bfloat16* Hb = reinterpret_cast<const __nv_bfloat16*>(H);
    int g2 = (M + 15) / 16;                     // 8 warps x 2 rows per CTA iteration
    if (g2 > cap) g2 = cap;
    if (G <= 2 && J <= 256) attn_logits_fwd_regs_kernel<1><<<g2, 256, 0, ST(stream)>>>(Hb, ldh, W2, b2, logits, M, J, G);
    else if (G <= 2 && J <= 512) attn_logits_fwd_regs_kernel<2><<<g2, 256, 0, ST(stream)>>>(Hb, ldh, W2, b2, logits, M, J, G);
    else if (G <= 2 && J <= 1024) attn_logits_fwd_regs_kernel<4><<<g2, 256, 0, ST(stream)>>>(Hb, ldh, W2, b2, logits, M, J, G);
    else attn_logits_fwd_kernel<true><<<grid, 256, smem, ST(stream)>>>(H, ldh, W2, b2, logits, M, J, G);
  } else {
    attn_logits_fwd_kernel<false><<<grid, 256, smem, ST(stream)>>>(H, ldh, W2, b2, logits, M, J, G);
  }
  VQA_LAUNCH_CHECK("attn_logits_fwd");
  return 0;
}

extern "C" int vqa_b200_attn_logits_bwd(const void* H, int h_dtype, int64_t ldh, const float* W2,
                                        const float* dlogits, void* dH, int dh_dtype, int64_t lddh,
                                        const float* out_scale, int rows_per_group, int relu_mask, float* dW2,
                                        float* db2, float* dbias_h, int M, int J, int G, void* stream) {
  if (!H || !W2 || !dlogits || !dH || !dW2 || M <= 0 || J <= 0 || G <= 0 || G > 4 || (J & 3) || J > 1024)
    return set_error(VQA_B200_EINVAL, "attn_logits_bwd: bad arguments (G 1..4, J a multiple of 4, J <= 1024)");
  if (rows_per_group <= 0) rows_per_group = 1;
  {
    const int hs = h_dtype == VQA_B200_BF16 ? 2 : 4, ds = dh_dtype == VQA_B200_BF16 ? 2 : 4;
    if ((reinterpret_cast<uintptr_t>(H) % (4 * hs)) || (reinterpret_cast<uintptr_t>(dH) % (4 * ds)) || (ldh % 4) || (lddh % 4))
      return set_error(VQA_B200_EALIGN, "attn_logits_bwd: H / dH rows must be aligned to 4 elements");
  }
  const bool hb = h_dtype == VQA_B200_BF16, db = dh_dtype == VQA_B200_BF16;
  // 128-bit accesses on both sides need bf16 H and dH with 16-byte aligned rows
  const bool wide = hb && db && (J % 8 == 0) && J <= 2048 && aligned16(H) && aligned16(dH) && (ldh % 8 == 0) && (lddh % 8 == 0);
  const int V = wide ? 8 : 4;
  const int blocks = sm_count() * 4;
  int rpb = (M + blocks - 1) / blocks;
  if (rpb < 8) rpb = 8;
  const int grid = (M + rpb - 1) / rpb;
  const int JT = (J + V - 1) / V, RLn = 256 / JT;
  const int GN = G <= 2 ? 2 : 4;
  const size_t smem_alb = (size_t)RLn * (GN + 1) * JT * V * sizeof(float);
#define LAUNCH_ALB(A_, B_, V_, G_)                                                                                  \
  attn_logits_bwd_kernel<A_, B_, V_, G_><<<grid, 256, smem_alb, ST(stream)>>>(H, ldh, W2, dlogits, dH, lddh, out_scale, \
                                                                           rows_per_group, relu_mask, dW2, db2,     \
                                                                           dbias_h, M, J, G, rpb)
#define LAUNCH_ALB_G(A_, B_, V_)            \
  do {                                      \
    if (GN == 2) LAUNCH_ALB(A_, B_, V_, 2); \
    else LAUNCH_ALB(A_, B_, V_, 4);         \
  } while (0)
  if (wide) LAUNCH_ALB_G(true, true, 8);
  else if (hb && db) LAUNCH_ALB_G(true, true, 4);
  else if (hb && !db) LAUNCH_ALB_G(true, false, 4);
  else if (!hb && db) LAUNCH_ALB_G(false, true, 4);
  else LAUNCH_ALB_G(false, false, 4);
#undef LAUNCH_ALB_G
#undef LAUNCH_ALB
  VQA_LAUNCH_CHECK("attn_logits_bwd");
  return 0;
}

static int pool_warps_per_cta(long long grid) {
  // resident CTAs per SM at the kernels' 62-72 registers per thread: 7 (4 warps), 14 (2 warps), 28+ (1 warp)
  const long long sms = sm_count();
  if (grid <= sms * 7) return 4;
  if (grid <= sms * 14) return 2;
  return 1;
}

extern "C" int vqa_b200_softmax_pool_fwd(const void* X, int x_dtype, const float* logits, float* att, float* pooled,
                                         int N, int L, int D, int G, int degenerate, void* stream) {
  if (!X || !logits || !pooled || N <= 0 || L <= 0 || D <= 0 || (G != 1 && G != 2))
    return set_error(VQA_B200_EINVAL, "softmax_pool_fwd: bad arguments (G must be 1 or 2)");
  const bool bf = x_dtype == VQA_B200_BF16;
  const int V = bf ? 8 : 4;
  if (D % V != 0 || !aligned16(X) || !aligned16(pooled))
    return set_error(VQA_B200_EALIGN, "softmax_pool_fwd: D must be a multiple of %d and X / pooled 16-byte aligned", V);
  const int chunks = (D + 32 * V - 1) / (32 * V);
  const long long grid = (long long)N * chunks;
  // warps per CTA: the largest of 4 / 2 / 1 for which the whole grid is one resident wave (32/NW CTAs per SM at the
  // kernels' 62 registers per thread); a ragged second wave costs more than the shorter per-warp row range gains
  const int nw = pool_warps_per_cta(grid);
  const size_t smem = ((size_t)G * L + (size_t)(nw - 1) * G * 32 * V) * sizeof(float);
  if (smem > 200 * 1024 || grid > 0x7fffffffLL) return set_error(VQA_B200_EINVAL, "softmax_pool_fwd: L too large");
#define LAUNCH_SPF_(B_, G_, W_)                                                                              \
  do {                                                                                                       \
    auto k = softmax_pool_fwd_kernel<B_, G_, W_>;                                                            \
    if (smem > 48 * 1024) VQA_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k<<<(int)grid, W_ * 32, smem, ST(stream)>>>(X, logits, att, pooled, L, D, chunks, degenerate);           \
  } while (0)
#define LAUNCH_SPF(B_, G_)                                                                                   \
  do {                                                                                                       \
    if (nw == 4) LAUNCH_SPF_(B_, G_, 4);                                                                     \
    else if (nw == 2) LAUNCH_SPF_(B_, G_, 2);                                                                \
    else LAUNCH_SPF_(B_, G_, 1);                                                                             \
  } while (0)
  if (bf && G == 2) LAUNCH_SPF(true, 2);
  else if (bf && G == 1) LAUNCH_SPF(true, 1);
  else if (!bf && G == 2) LAUNCH_SPF(false, 2);
  else LAUNCH_SPF(false, 1);
#undef LAUNCH_SPF
#undef LAUNCH_SPF_
  VQA_LAUNCH_CHECK("softmax_pool_fwd");
  return 0;
}

extern "C" int vqa_b200_softmax_pool_bwd(const void* X, int x_dtype, const float* att, const float* dpooled,
                                         const float* datt_extra, float* dlogits, uint32_t* done, float* dX, int N,
                                         int L, int D, int G, int degenerate, int accumulate_dx, void* stream) {
  if (!X || !att || !dpooled || !dlogits || !done || N <= 0 || L <= 0 || D <= 0 || (G != 1 && G != 2))
    return set_error(VQA_B200_EINVAL, "softmax_pool_bwd: bad arguments (G must be 1 or 2)");
  const bool bf = x_dtype == VQA_B200_BF16;
  const int V = bf ? 8 : 4;
  if (D % V != 0 || D % 4 != 0 || !aligned16(X) || !aligned16(dpooled) || (dX && !aligned16(dX)))
    return set_error(VQA_B200_EALIGN, "softmax_pool_bwd: D must be a multiple of %d, operands 16-byte aligned", V);
  const int chunks = (D + 32 * V - 1) / (32 * V);
  const long long grid = (long long)N * chunks;
  const size_t smem2 = (2 * (size_t)G * L + G) * sizeof(float);
  if (smem2 > 200 * 1024 || grid > 0x7fffffffLL) return set_error(VQA_B200_EINVAL, "softmax_pool_bwd: L too large");
  // datt accumulators and the per-sample completion tickets start at zero: ONE memset when the caller laid them out
  // back to back (the Python operator does), else two
  if (reinterpret_cast<const float*>(done) == dlogits + (size_t)N * G * L) {
    VQA_CUDA_CHECK(cudaMemsetAsync(dlogits, 0, ((size_t)N * G * L + N) * sizeof(float), ST(stream)));
  } else {
    VQA_CUDA_CHECK(cudaMemsetAsync(dlogits, 0, (size_t)N * G * L * sizeof(float), ST(stream)));
    VQA_CUDA_CHECK(cudaMemsetAsync(done, 0, (size_t)N * sizeof(uint32_t), ST(stream)));
  }
  const int nw = pool_warps_per_cta(grid);
#define LAUNCH_SPB__(B_, G_, X_, W_)                                                                         \
  do {                                                                                                       \
    auto k = softmax_pool_bwd_kernel<B_, G_, X_, W_>;                                                        \
    if (smem2 > 48 * 1024)                                                                                   \
      VQA_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));      \
    k<<<(int)grid, W_ * 32, smem2, ST(stream)>>>(X, att, dpooled, dlogits, dX, datt_extra, done, L, D, chunks, \
                                                 degenerate, accumulate_dx);                                 \
  } while (0)
#define LAUNCH_SPB_(B_, G_, X_)                                                                              \
  do {                                                                                                       \
    if (nw == 4) LAUNCH_SPB__(B_, G_, X_, 4);                                                                \
    else if (nw == 2) LAUNCH_SPB__(B_, G_, X_, 2);                                                           \
    else LAUNCH_SPB__(B_, G_, X_, 1);                                                                        \
  } while (0)
#define LAUNCH_SPB(B_, G_)                                                                                   \
  do {                                                                                                       \
    if (dX != nullptr) LAUNCH_SPB_(B_, G_, true);                                                            \
    else LAUNCH_SPB_(B_, G_, false);                                                                         \
  } while (0)
  if (bf && G == 2) LAUNCH_SPB(true, 2);
  else if (bf && G == 1) LAUNCH_SPB(true, 1);
  else if (!bf && G == 2) LAUNCH_SPB(false, 2);
  else LAUNCH_SPB(false, 1);
#undef LAUNCH_SPB
#undef LAUNCH_SPB_
#undef LAUNCH_SPB__
  VQA_LAUNCH_CHECK("softmax_pool_bwd");
  return 0;
}

extern "C" int vqa_b200_mfb_bwd(const void* G, int g_dtype, int64_t ldg, const void* Y, int y_dtype, int64_t ldy,
                                const float* inv, const float* t, const float* Q, int64_t ldq, const void* keep,
                                int keep_dtype, void* dI, int di_dtype, float* dQ, float* dbias, int rows_per_group,
                                int M, int N, int seg_cols, const float* extra, const float* dprod_in, float* dExtra,
                                float drop_p, uint32_t seed, const uint32_t* seed_dev, void* stream) {
  if (!G || !Y || !inv || !t || !Q || !keep || !dI || !dQ || M <= 0 || N <= 0 || N % 20 != 0)
    return set_error(VQA_B200_EINVAL, "mfb_bwd: bad arguments (N %% 20 == 0 required)");
  if (seg_cols <= 0) seg_cols = N;
  if (N % seg_cols != 0 || seg_cols % 20 != 0)
    return set_error(VQA_B200_EINVAL, "mfb_bwd: seg_cols (%d) must divide N (%d) and be a multiple of 20", seg_cols, N);
  if (rows_per_group <= 0) rows_per_group = 1;
  if (g_dtype != y_dtype || keep_dtype != di_dtype)
    return set_error(VQA_B200_EINVAL, "mfb_bwd: g/y and keep/dI must share a dtype (g=%d y=%d keep=%d dI=%d)", g_dtype,
                     y_dtype, keep_dtype, di_dtype);
  const int ygs = g_dtype == VQA_B200_BF16 ? 2 : 4;
  if (!aligned16(Q) || (ldq * 4) % 16 != 0 || !aligned16(keep) || !aligned16(dI) || !aligned16(dQ) || !aligned16(G) ||
      !aligned16(Y) || (ldg * ygs) % (2 * ygs) != 0 || (ldy * ygs) % (2 * ygs) != 0 || (dbias && !aligned16(dbias)))
    return set_error(VQA_B200_EALIGN, "mfb_bwd: operands must be 16-byte aligned, g / y row pitches multiples of 4 elements");
  const int col_blocks = (N / 20 + 255) / 256;
  if ((extra || dExtra) && rows_per_group != 1)
    return set_error(VQA_B200_EINVAL, "mfb_bwd: the cascade multiplier is defined for vector blocks (rows_per_group == 1)");
  if ((extra && !aligned16(extra)) || (dprod_in && !aligned16(dprod_in)) || (dExtra && !aligned16(dExtra)))
    return set_error(VQA_B200_EALIGN, "mfb_bwd: extra / dprod_in / dExtra must be 16-byte aligned");
  uint32_t th; float sc;
  drop_params(drop_p, &th, &sc);
  const int groups = (M + rows_per_group - 1) / rows_per_group;
  const bool yb = y_dtype == VQA_B200_BF16, kb = keep_dtype == VQA_B200_BF16;
  const bool cascade = seg_cols != N || extra != nullptr || dprod_in != nullptr || dExtra != nullptr;
  // enough CTAs to fill the machine: slice the rows of a group when there are few groups.  (Measured on the grid MFB,
  // 256 groups of 196 rows, two resident CTAs per SM: 4 slices 0.271 ms, 5: 0.278, 8: 0.283, 14: 0.307 -- the prefetch
  // ring's fill / drain per CTA outweighs the wave quantisation, so the slices stay as long as they can.)
  int slices = (sm_count() * 6 + groups - 1) / groups;
  if (slices > rows_per_group) slices = rows_per_group;
  if (slices < 1) slices = 1;
  const int rps = (rows_per_group + slices - 1) / slices;
  slices = (rows_per_group + rps - 1) / rps;
  if (slices > 1) VQA_CUDA_CHECK(cudaMemsetAsync(dQ, 0, (size_t)groups * N * sizeof(float), ST(stream)));
  dim3 grid(groups, slices, col_blocks);
#define LAUNCH_MB(A_, B_)                                                                                        \
  do {                                                                                                           \
    const size_t smem = (size_t)MFB_BWD_PREFETCH * 256 * (5 * ((B_) ? 8 : 16) + 2 * ((A_) ? 8 : 16));            \
    auto k = cascade ? mfb_bwd_kernel<A_, B_, true> : mfb_bwd_kernel<A_, B_, false>;                             \
    VQA_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    k<<<grid, 256, smem, ST(stream)>>>(G, ldg, Y, ldy, inv, t, Q, ldq, keep, dI, dQ, dbias, extra, dprod_in,    \
                                       dExtra, rows_per_group, rps, M, N, seg_cols, seed, seed_dev, th, sc);     \
  } while (0)
  if (yb && kb) LAUNCH_MB(true, true);
  else if (yb && !kb) LAUNCH_MB(true, false);
  else if (!yb && kb) LAUNCH_MB(false, true);
  else LAUNCH_MB(false, false);
#undef LAUNCH_MB
  VQA_LAUNCH_CHECK("mfb_bwd");
  return 0;
}

extern "C" int vqa_b200_norm_bwd_prep(const float* d, int64_t ldd, const void* Y, int y_dtype, int64_t ldy,
                                      const float* inv, float* g, int64_t ldg, float* t, int rows_per_group, int M,
                                      int No, void* stream) {
  if (!d || !Y || !inv || !g || !t || M <= 0 || No <= 0) return set_error(VQA_B200_EINVAL, "norm_bwd_prep: bad arguments");
  if (rows_per_group <= 0) rows_per_group = 1;
  int grid = (M + 7) / 8;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  if (y_dtype == VQA_B200_BF16)
    norm_bwd_prep_kernel<true><<<grid, 256, 0, ST(stream)>>>(d, ldd, Y, ldy, inv, g, ldg, t, rows_per_group, M, No);
  else
    norm_bwd_prep_kernel<false><<<grid, 256, 0, ST(stream)>>>(d, ldd, Y, ldy, inv, g, ldg, t, rows_per_group, M, No);
  VQA_LAUNCH_CHECK("norm_bwd_prep");
  return 0;
}

extern "C" int vqa_b200_inv_norm(const float* ssq, float* inv, int n, void* stream) {
  if (!ssq || !inv || n <= 0) return set_error(VQA_B200_EINVAL, "inv_norm: bad arguments");
  inv_norm_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(ssq, inv, n);
  VQA_LAUNCH_CHECK("inv_norm");
  return 0;
}

extern "C" int vqa_b200_scale_rows(const void* Y, int y_dtype, int64_t ldy, const float* inv, int rows_per_group,
                                   float* out, int64_t ldo, int M, int No, void* stream) {
  if (!Y || !inv || !out || M <= 0 || No <= 0) return set_error(VQA_B200_EINVAL, "scale_rows: bad arguments");
  if (rows_per_group <= 0) rows_per_group = 1;
  const int grid = ew_grid((long long)M * No, 256);
  if (y_dtype == VQA_B200_BF16)
    scale_rows_kernel<true><<<grid, 256, 0, ST(stream)>>>(Y, ldy, inv, rows_per_group, out, ldo, M, No);
  else
    scale_rows_kernel<false><<<grid, 256, 0, ST(stream)>>>(Y, ldy, inv, rows_per_group, out, ldo, M, No);
  VQA_LAUNCH_CHECK("scale_rows");
  return 0;
}

static void strip_grid(int M, int J, dim3* grid, int* rpb) {
  const int col_blocks = (J + 63) / 64;
  int strips = (sm_count() * 8 + col_blocks - 1) / col_blocks;
  if (strips > (M + 3) / 4) strips = (M + 3) / 4;
  if (strips < 1) strips = 1;
  *rpb = (M + strips - 1) / strips;
  *grid = dim3(col_blocks, (M + *rpb - 1) / *rpb);
}

extern "C" int vqa_b200_group_dot(const void* A, int a_dtype, int64_t lda, const void* B, int b_dtype, int64_t ldb,
                                  float* t, int rows_per_group, int M, int No, void* stream) {
  if (!A || !B || !t || M <= 0 || No <= 0) return set_error(VQA_B200_EINVAL, "group_dot: bad arguments");
  if (rows_per_group <= 0) rows_per_group = 1;
  int grid = (M + 7) / 8;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  group_dot_kernel<<<grid, 256, 0, ST(stream)>>>(A, a_dtype == VQA_B200_BF16, lda, B, b_dtype == VQA_B200_BF16, ldb, t,
                                                 rows_per_group, M, No);
  VQA_LAUNCH_CHECK("group_dot");
  return 0;
}

extern "C" int vqa_b200_colsum(const void* X, int x_dtype, int64_t ldx, float* out, int M, int J, void* stream) {
  if (!X || !out || M <= 0 || J <= 0) return set_error(VQA_B200_EINVAL, "colsum: bad arguments");
  const bool bf = x_dtype == VQA_B200_BF16;
  const int V = bf ? 8 : 4;
  if (J % V == 0 && ldx % V == 0 && aligned16(X)) {
    const int col_blocks = (J + 32 * V - 1) / (32 * V);
    int strips = (sm_count() * 4 + col_blocks - 1) / col_blocks;
    if (strips > (M + 31) / 32) strips = (M + 31) / 32;
    if (strips < 1) strips = 1;
    const int rpb = (M + strips - 1) / strips;
    const dim3 grid(col_blocks, (M + rpb - 1) / rpb);
    if (bf) colsum_vec_kernel<true><<<grid, 256, 0, ST(stream)>>>(X, ldx, out, M, J, rpb);
    else colsum_vec_kernel<false><<<grid, 256, 0, ST(stream)>>>(X, ldx, out, M, J, rpb);
    VQA_LAUNCH_CHECK("colsum");
    return 0;
  }
  dim3 grid; int rpb;
  strip_grid(M, J, &grid, &rpb);
  colsum_kernel<<<grid, 256, 0, ST(stream)>>>(X, bf, ldx, out, M, J, rpb);
  VQA_LAUNCH_CHECK("colsum");
  return 0;
}

extern "C" int vqa_b200_relu_bwd(const void* D, int d_dtype, int64_t ldd, const void* H, int h_dtype, int64_t ldh,
                                 void* out, int o_dtype, int64_t ldo, const float* scale, int rows_per_group,
                                 float* dbias, int M, int J, void* stream) {
  if (!D || !H || !out || M <= 0 || J <= 0) return set_error(VQA_B200_EINVAL, "relu_bwd: bad arguments");
  if (rows_per_group <= 0) rows_per_group = 1;
  dim3 grid; int rpb;
  strip_grid(M, J, &grid, &rpb);
  relu_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(D, d_dtype == VQA_B200_BF16, ldd, H, h_dtype == VQA_B200_BF16, ldh, out,
                                                o_dtype == VQA_B200_BF16, ldo, scale, rows_per_group, dbias, M, J, rpb);
  VQA_LAUNCH_CHECK("relu_bwd");
  return 0;
}


// =====================================================================================
// Elementwise steps of HieCoAtten / modules.py (hieCoAtten.py:25-50, modules.py:26-33,103-109)
// =====================================================================================
namespace vqa {

// out = dropout(act(x (+ add) (+ bias[col])))   act: 0 none, 1 relu, 2 tanh, 3 sigmoid;  mask hash on (row, col)
__global__ void act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ add,
                               const float* __restrict__ bias, float* __restrict__ out, long long rows, int cols,
                               int act, uint32_t seed, const uint32_t* __restrict__ seed_dev, uint32_t thresh16,
                               float scale) {
  seed = effective_seed(seed, seed_dev);
  const long long total = rows * cols;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int c = (int)(i % cols);
    const long long r = i / cols;
    float v = x[i];
    if (add) v += add[i];
    if (bias) v += bias[c];
    if (act == 1) v = fmaxf(v, 0.f);
    else if (act == 2) v = tanhf(v);
    else if (act == 3) v = 1.f / (1.f + expf(-v));
    if (thresh16) v = dropout_keep(seed, (uint32_t)r, (uint32_t)c, thresh16) ? v * scale : 0.f;
    out[i] = v;
  }
}

// dpre = dout * mask_scale * act'(.)  with the activation derivative recovered from the SAVED OUTPUT h
// (h = act(pre) * mask_scale):  relu: h > 0;  tanh: 1 - (h / scale)^2;  none: 1.   dbias[col] += dpre.
__global__ void __launch_bounds__(256) act_bwd_kernel(const void* __restrict__ D, int dbf, long long ldd,
                                                      const void* __restrict__ H, int hbf, long long ldh,
                                                      void* __restrict__ out, int obf, long long ldo,
                                                      float* __restrict__ dbias, int M, int J, int rows_per_block,
                                                      int act, uint32_t seed, const uint32_t* __restrict__ seed_dev, uint32_t thresh16,
                               float scale) {
  seed = effective_seed(seed, seed_dev);
  __shared__ float red[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int j = blockIdx.x * 64 + tx;
  const int m0 = blockIdx.y * rows_per_block, m1 = min(M, m0 + rows_per_block);
  float acc = 0.f;
  if (j < J) {
    for (int m = m0 + ty; m < m1; m += 4) {
      float d = ld_any(D, dbf, (long long)m * ldd + j);
      const float h = ld_any(H, hbf, (long long)m * ldh + j);
      if (thresh16) d = dropout_keep(seed, (uint32_t)m, (uint32_t)j, thresh16) ? d * scale : 0.f;
      if (act == 1) { if (!(h > 0.f)) d = 0.f; }
      else if (act == 2) { const float y = h / scale; d *= (1.f - y * y); }
      acc += d;
      st_any(out, obf, (long long)m * ldo + j, d);
    }
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (dbias && ty == 0 && j < J) atomicAdd(dbias + j, red[0][tx] + red[1][tx] + red[2][tx] + red[3][tx]);
}

// row softmax over the last axis (modules.py:90): warp per row
__global__ void __launch_bounds__(256) row_softmax_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * warps) {
    const float* xr = x + r * cols;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, xr[c]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += __expf(xr[c] - mx);
    s = warp_sum(s);
    const float inv = 1.f / s;
    for (int c = lane; c < cols; c += 32) y[r * cols + c] = __expf(xr[c] - mx) * inv;
  }
}
// dx = y * (dy - sum(y * dy))
__global__ void __launch_bounds__(256) row_softmax_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                              float* __restrict__ dx, long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  for (long long r = (long long)blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += (long long)gridDim.x * warps) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += y[r * cols + c] * dy[r * cols + c];
    s = warp_sum(s);
    for (int c = lane; c < cols; c += 32) dx[r * cols + c] = y[r * cols + c] * (dy[r * cols + c] - s);
  }
}

// gated tanh (modules.py:103-109): o = tanh(a) * sigmoid(b);  backward: da = do * sig(b) * (1 - tanh(a)^2), db = do * tanh(a) * sig(b)(1 - sig(b))
__global__ void gate_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    o[i] = tanhf(a[i]) * (1.f / (1.f + expf(-b[i])));
}
__global__ void gate_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ d_o,
                                float* __restrict__ da, float* __restrict__ db, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float t = tanhf(a[i]);
    const float s = 1.f / (1.f + expf(-b[i]));
    da[i] = d_o[i] * s * (1.f - t * t);
    db[i] = d_o[i] * t * s * (1.f - s);
  }
}

}  // namespace vqa

extern "C" int vqa_b200_act_fwd(const float* x, const float* add, const float* bias, float* out, int64_t rows, int cols,
                                int act, float drop_p, uint32_t seed, const uint32_t* seed_dev, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0) return set_error(VQA_B200_EINVAL, "act_fwd: bad arguments");
  uint32_t th; float sc;
  drop_params(drop_p, &th, &sc);
  act_fwd_kernel<<<ew_grid(rows * cols, 256), 256, 0, ST(stream)>>>(x, add, bias, out, rows, cols, act, seed, seed_dev,
                                                                    th, sc);
  VQA_LAUNCH_CHECK("act_fwd");
  return 0;
}

extern "C" int vqa_b200_act_bwd(const void* D, int d_dtype, int64_t ldd, const void* H, int h_dtype, int64_t ldh,
                                void* out, int o_dtype, int64_t ldo, float* dbias, int M, int J, int act, float drop_p,
                                uint32_t seed, const uint32_t* seed_dev, void* stream) {
  if (!D || !H || !out || M <= 0 || J <= 0) return set_error(VQA_B200_EINVAL, "act_bwd: bad arguments");
  uint32_t th; float sc;
  drop_params(drop_p, &th, &sc);
  dim3 grid; int rpb;
  strip_grid(M, J, &grid, &rpb);
  act_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(D, d_dtype == VQA_B200_BF16, ldd, H, h_dtype == VQA_B200_BF16, ldh, out,
                                               o_dtype == VQA_B200_BF16, ldo, dbias, M, J, rpb, act, seed, seed_dev,
                                               th, sc);
  VQA_LAUNCH_CHECK("act_bwd");
  return 0;
}

extern "C" int vqa_b200_row_softmax_fwd(const float* x, float* y, int64_t rows, int cols, void* stream) {
  if (!x || !y || rows <= 0 || cols <= 0) return set_error(VQA_B200_EINVAL, "row_softmax_fwd: bad arguments");
  long long grid = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  row_softmax_fwd_kernel<<<(int)grid, 256, 0, ST(stream)>>>(x, y, rows, cols);
  VQA_LAUNCH_CHECK("row_softmax_fwd");
  return 0;
}

extern "C" int vqa_b200_row_softmax_bwd(const float* y, const float* dy, float* dx, int64_t rows, int cols, void* stream) {
  if (!y || !dy || !dx || rows <= 0 || cols <= 0) return set_error(VQA_B200_EINVAL, "row_softmax_bwd: bad arguments");
  long long grid = (rows + 7) / 8;
  const long long cap = (long long)sm_count() * 8;
  if (grid > cap) grid = cap;
  row_softmax_bwd_kernel<<<(int)grid, 256, 0, ST(stream)>>>(y, dy, dx, rows, cols);
  VQA_LAUNCH_CHECK("row_softmax_bwd");
  return 0;
}

extern "C" int vqa_b200_gate_fwd(const float* a, const float* b, float* o, int64_t n, void* stream) {
  if (!a || !b || !o || n <= 0) return set_error(VQA_B200_EINVAL, "gate_fwd: bad arguments");
  gate_fwd_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(a, b, o, n);
  VQA_LAUNCH_CHECK("gate_fwd");
  return 0;
}

extern "C" int vqa_b200_gate_bwd(const float* a, const float* b, const float* d_o, float* da, float* db, int64_t n,
                                 void* stream) {
  if (!a || !b || !d_o || !da || !db || n <= 0) return set_error(VQA_B200_EINVAL, "gate_bwd: bad arguments");
  gate_bwd_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(a, b, d_o, da, db, n);
  VQA_LAUNCH_CHECK("gate_bwd");
  return 0;
}

// =====================================================================================
// classifier tail (mhb_coAtt.py:147-151 + solver.py:148-153): log-softmax over the answers + argmax in one pass set
// =====================================================================================
namespace vqa {
namespace {
// one warp per row: max / argmax, sum of exponentials, then the log-probabilities (the row -- 12 KB for 3000 answers --
// stays in L1 between the three sweeps, so HBM sees one read and one write of the logits)
__global__ void __launch_bounds__(256) logsoftmax_argmax_kernel(const float* __restrict__ logits, long long ldl,
                                                                float* __restrict__ logp, long long ldo,
                                                                long long* __restrict__ pred,
                                                                float* __restrict__ pred_logp, int M, int N) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* x = logits + (long long)row * ldl;
  float mx = -INFINITY;
  int am = 0;
  for (int c = lane; c < N; c += 32) {
    const float v = x[c];
    if (v > mx) { mx = v; am = c; }               // first maximum wins inside a lane ...
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (ov > mx || (ov == mx && oa < am)) { mx = ov; am = oa; }     // ... and across lanes (torch.argmax: lowest index)
  }
  float se = 0.f;
  for (int c = lane; c < N; c += 32) se += __expf(x[c] - mx);
  se = warp_sum(se);
  const float lse = mx + __logf(se);
  if (logp != nullptr) {
    float* y = logp + (long long)row * ldo;
    for (int c = lane; c < N; c += 32) y[c] = x[c] - lse;
  }
  if (lane == 0) {
    if (pred) pred[row] = am;
    if (pred_logp) pred_logp[row] = mx - lse;
  }
}
}  // namespace
}  // namespace vqa

extern "C" int vqa_b200_logsoftmax_argmax(const float* logits, int64_t ldl, float* logp, int64_t ldo, int64_t* pred,
                                          float* pred_logp, int M, int N, void* stream) {
  if (!logits || M <= 0 || N <= 0 || (!logp && !pred && !pred_logp))
    return set_error(VQA_B200_EINVAL, "logsoftmax_argmax: bad arguments");
  vqa::logsoftmax_argmax_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(logits, ldl, logp, ldo, (long long*)pred, pred_logp,
                                                                     M, N);
  VQA_LAUNCH_CHECK("logsoftmax_argmax");
  return 0;
}

// =====================================================================================
// Training loss of the solver (solver.py:26-29,77-92 with the soft answers of utils.py:250-265):
//   KLDivLoss(reduction='mean')(log_softmax(logits, 1), target) = sum_{m,n} (xlogy(t, t) - t * logp) / (M * N)
// ATen runs it as ten launches (log-softmax, mul, xlogy, sub, mean, and their backwards) over [M, 3000]; here one pass
// per direction, one warp per row.  The forward leaves the row statistics the backward needs (log-sum-exp and sum_n t):
//   dL/dlogits[m,n] = g * (softmax[m,n] * sum_n t[m,:] - t[m,n]) / (M * N)
// =====================================================================================
namespace vqa {
namespace {
__global__ void __launch_bounds__(256) kldiv_logsoftmax_fwd_kernel(const float* __restrict__ logits, long long ldl,
                                                                   const float* __restrict__ target, long long ldt,
                                                                   float* __restrict__ loss, float* __restrict__ lse_out,
                                                                   float* __restrict__ tsum_out, int M, int N,
                                                                   float scale) {
  __shared__ float part[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + w;
  float rl = 0.f;
  if (row < M) {
    const float* x = logits + (long long)row * ldl;
    const float* t = target + (long long)row * ldt;
    float mx = -INFINITY;
    for (int c = lane; c < N; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < N; c += 32) se += __expf(x[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    float ts = 0.f;
    for (int c = lane; c < N; c += 32) {
      const float tv = __ldg(t + c);
      ts += tv;
      if (tv > 0.f) rl += tv * (__logf(tv) - (x[c] - lse));      // xlogy(t, t) - t * logp;  t == 0 contributes nothing
      else if (tv < 0.f) rl = NAN;                                // as torch: xlogy of a negative target is NaN
    }
    ts = warp_sum(ts);
    rl = warp_sum(rl);
    if (lane == 0) {
      lse_out[row] = lse;
      tsum_out[row] = ts;
    }
  }
  if (lane == 0) part[w] = rl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float sres = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) sres += part[i];
    atomicAdd(loss, sres * scale);
  }
}

__global__ void __launch_bounds__(256) kldiv_logsoftmax_bwd_kernel(const float* __restrict__ logits, long long ldl,
                                                                   const float* __restrict__ target, long long ldt,
                                                                   const float* __restrict__ lse, const float* __restrict__ tsum,
                                                                   const float* __restrict__ gout, float* __restrict__ dlogits,
                                                                   long long ldd, int M, int N, float scale) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float g = (gout ? __ldg(gout) : 1.f) * scale;
  const float l = lse[row], ts = tsum[row];
  const float* x = logits + (long long)row * ldl;
  const float* t = target + (long long)row * ldt;
  float* d = dlogits + (long long)row * ldd;
  for (int c = lane; c < N; c += 32) d[c] = g * (__expf(x[c] - l) * ts - __ldg(t + c));
}
}  // namespace
}  // namespace vqa

extern "C" int vqa_b200_kldiv_logsoftmax_fwd(const float* logits, int64_t ldl, const float* target, int64_t ldt,
                                             float* loss, float* lse, float* tsum, int M, int N, void* stream) {
  if (!logits || !target || !loss || !lse || !tsum || M <= 0 || N <= 0 || ldl < N || ldt < N)
    return set_error(VQA_B200_EINVAL, "kldiv_logsoftmax_fwd: bad arguments");
  vqa::kldiv_logsoftmax_fwd_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(logits, ldl, target, ldt, loss, lse, tsum, M, N,
                                                                        1.0f / ((float)M * (float)N));
  VQA_LAUNCH_CHECK("kldiv_logsoftmax_fwd");
  return 0;
}

extern "C" int vqa_b200_kldiv_logsoftmax_bwd(const float* logits, int64_t ldl, const float* target, int64_t ldt,
                                             const float* lse, const float* tsum, const float* gout, float* dlogits,
                                             int64_t ldd, int M, int N, void* stream) {
  if (!logits || !target || !lse || !tsum || !dlogits || M <= 0 || N <= 0 || ldl < N || ldt < N || ldd < N)
    return set_error(VQA_B200_EINVAL, "kldiv_logsoftmax_bwd: bad arguments");
  vqa::kldiv_logsoftmax_bwd_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(logits, ldl, target, ldt, lse, tsum, gout, dlogits,
                                                                        ldd, M, N, 1.0f / ((float)M * (float)N));
  VQA_LAUNCH_CHECK("kldiv_logsoftmax_bwd");
  return 0;
}
