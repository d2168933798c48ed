// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[m, n] = epilogue( sum_k A(m, k) * B(n, k) )        bf16 operands, fp32 accumulation in TMEM
//
// Operands are fetched by TMA (SWIZZLE_128B) into a multi-stage shared-memory ring and consumed
// by single-thread tcgen05.mma (cta_group::1, M = 128, N = BN, K = 16).  Each operand may be
//   * K-major  : memory [rows, K], K contiguous     (forward Linear / 1x1-conv: x and W)
//   * MN-major : memory [K, rows], rows contiguous  (dgrad: W as B;  wgrad: dY^T as A and X^T as B)
// so forward, dgrad and wgrad of every Linear / 1x1 conv on the path run on the same kernel with no
// transposed copies.  The accumulator is double-buffered in TMEM (2 x BN columns) so the epilogue of
// tile i overlaps the main loop of tile i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue.  Warp w reads TMEM lanes 32*(w%4) .. +31 (thread == accumulator row); the two warps of a
// lane quadrant take alternating column chunks, so every SM sub-partition has two epilogue warps to hide the
// tcgen05.ld / shared / L1 latencies of the fused epilogues behind each other.
#pragma once
#include "ptx.cuh"

// Pipeline wait-cycle counters (where do the producer / MMA threads wait?) exist only in -DVQA_B200_DEBUG builds; release
// kernels carry no clock64() reads.
#ifdef VQA_B200_DEBUG
#define VQA_DBG(...) __VA_ARGS__
#else
#define VQA_DBG(...)
#endif

namespace vqa {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;       // 64 bf16 = 128 bytes = one SWIZZLE_128B row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps (two per TMEM lane quadrant / SMSP)

enum EpiKind { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_MFB = 2 };

struct GemmArgs {
  int M, N, K;                          // per batch entry
  int batch;                            // independent [M,N,K] problems (hieCoAtten.py:32,38,45 bmm); 1 = plain GEMM
  int m_blocks, n_blocks, k_blocks, k_split;
  unsigned long long* dbg;              // optional [8] cycle counters of CTA 0 (debug: where the pipeline waits)
  int full_units;                       // leading tiles that are NOT split along K (tail-wave split, see unit_decode)
  int a_mn, b_mn;                       // operand majorness (0 = K-major, 1 = MN-major)
  uint64_t a_desc_hi, b_desc_hi;        // smem descriptor without the start address
  uint32_t a_kadv, b_kadv;              // start-address advance (16-byte units) per UMMA_K step
  uint32_t idesc;
  // ---- EPI_STORE / EPI_ATOMIC
  void* C;                              // [batch, M, N] row-major, ldc elements per row
  long long ldc;
  long long c_bstride;                  // elements between batch entries of C (and of `add`, `dot_with`)
  int c_bf16;                           // 1: bf16 output, 0: fp32
  int vec_ok;                           // rows are 16-byte aligned -> vector stores allowed
  const float* bias;                    // [N] or null
  const float* row_scale;               // [ceil(M / rows_per_group)] or null: out = acc * scale[m / rpg] + bias
  int rows_per_group;
  int act;                              // 0 none, 1 ReLU, 2 tanh
  const void* add;                      // optional addend [batch, M, N] (same ld / batch stride as C), before act
  int add_bf16;
  uint32_t st_drop_seed, st_drop_thresh16;   // EPI_STORE dropout after the activation (F.dropout, hieCoAtten.py:26-46)
  float st_drop_scale;
  const uint32_t* seed_dev;             // optional device-side salt of both dropout seeds (CUDA-graph replay), see ptx.cuh
  const __nv_bfloat16* dot_with;        // optional [M, N] (ld_dot): dot_out[m / rpg] += sum_n out * dot_with
  long long ld_dot;
  float* dot_out;
  // ---- EPI_MFB (mhb_coAtt.py:94-106): acc = image projection, columns c = 5*o + j
  const float* mfb_q;                   // [groups, N] projected question vector (bias included), ld = mfb_ldq
  long long mfb_ldq;
  const float* mfb_extra;               // optional [groups, N] (ld = mfb_ldq): a second multiplier of the product -- the
                                        // dropped-out product of the previous block in MHB's cascade (mhb_coAtt.py:204-205)
  float* mfb_prod;                      // optional fp32 [M, N] (ld = N): the dropped-out product (acc+b)*mask*Q*extra itself,
                                        // i.e. what the next block of the cascade multiplies by
  void* mfb_y;                          // [M, N/5] signed-sqrt of the k-pooled product (bf16 or fp32)
  long long mfb_ldy;
  int mfb_y_bf16;
  float* mfb_ssq;                       // [groups, nseg] += sum |z|  (== sum y^2, for the per-sample L2 norm)
  int mfb_seg_cols;                     // columns per L2-norm segment (N = one segment; 5000 when two MFB blocks share
                                        // one launch: [img_proj2; img_proj3] -> N = 10000, mhb_coAtt.py:125,137)
  void* mfb_keep;                       // optional [M, N] (ld = N): (acc + bias) * mask, saved for backward
  int mfb_keep_f32;                     // keep dtype: 0 = bf16, 1 = fp32
  uint32_t drop_seed, drop_thresh16;    // thresh16 == 0 -> no dropout
  float drop_scale;
};

template <int BN, bool CTA2 = false>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * 128;
  static constexpr int BN_CTA = CTA2 ? BN / 2 : BN;             // B rows held by one CTA (pairs split the B tile)
  static constexpr int B_BYTES = BN_CTA * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CTA2 ? 6 : ((BN <= 128) ? 6 : 4);
  static constexpr int ACC_STRIDE = (BN <= 128) ? 128 : 256;   // TMEM columns per accumulator stage
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  // EPI_MFB (BN == 240) stages the bf16 `keep` tile through smem for coalesced stores: per epilogue warp
  // 32 rows x 40 columns
  static constexpr int KEEP_PITCH = 80;                 // 40 bf16 columns per row: 20-word pitch -> conflict-free 16-byte reads
  static constexpr int KEEP_STAGE_BYTES = (BN == 240) ? 8 * 32 * KEEP_PITCH : 0;
  // EPI_STORE / EPI_ATOMIC transpose their 32-row x 32-column register chunks through shared memory (XOR-swizzled 16-byte
  // pieces, 4 KB per epilogue warp) so that every global store / reduction instruction covers whole 128-byte (fp32) or
  // 64-byte (bf16) row segments instead of 32 pieces of 16 bytes in 32 different rows
  static constexpr int OUT_STAGE_BYTES = (BN == 240) ? 0 : 8 * 32 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + KEEP_STAGE_BYTES + OUT_STAGE_BYTES + 1024;   // + slack
};

// Thread `lane` (= row of a 32-row chunk) parks its 32 fp32 values as eight 16-byte pieces, piece j at slot j ^ (row & 7):
// conflict-free for the row-wise writes (quarter-warps hit eight different bank groups) and for the read-back, where
// eight consecutive lanes fetch the eight pieces of ONE row.
__device__ __forceinline__ void stage_rows_f32(uint8_t* tile, int lane, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<float4*>(tile + lane * 128 + ((q ^ (lane & 7)) << 4)) =
        make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
}

// Work units of the persistent grid.  Units [0, full_units) are whole tiles; every remaining tile is cut into
// k_split slices along K.  With full_units = floor(tiles / #CTAs) * #CTAs only the ragged last wave is split, so the
// grid stays in K-lockstep (L2 reuse of the operand panels) and the quantisation loss disappears (accumulate mode).
struct Unit { int m_blk, n_blk, bz, kb0, kb1; };
__device__ __forceinline__ Unit unit_decode(const GemmArgs& p, int u) {
  int tile, kb0 = 0, kb1 = p.k_blocks;
  if (u < p.full_units) {
    tile = u;
  } else {
    const int v = u - p.full_units;
    tile = p.full_units + v / p.k_split;
    const int ks = v % p.k_split;
    kb0 = (int)(((long long)ks * p.k_blocks) / p.k_split);
    kb1 = (int)(((long long)(ks + 1) * p.k_blocks) / p.k_split);
  }
  Unit r;
  r.n_blk = tile % p.n_blocks;
  const int r2 = tile / p.n_blocks;
  r.m_blk = r2 % p.m_blocks;
  r.bz = r2 / p.m_blocks;
  r.kb0 = kb0; r.kb1 = kb1;
  return r;
}

// MFBX (EPI_MFB only): the vector-block extras -- several L2-norm segments per row (two MFB blocks in one launch) and
// MHB's cascade multiplier / product output.  The grid MFB (img_conv1d, 91 % of the model's FLOPs) is compiled without
// them: its epilogue, not the MMA, bounds that kernel, and every extra test in its inner loop shows (0.77 -> 0.82 ms).
template <int BN, int EPI, bool CTA2 = false, bool MFBX = false>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmArgs p) {
  using Cfg = GemmCfg<BN, CTA2>;
  // CTA pair (cta_group::2): launched as clusters of 2; both CTAs run the producer and the epilogue on their own
  // 128 rows, the leader (rank 0) issues the MMAs for the whole 256-row tile.
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const int grid_units = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;     // persistent workers (CTAs or pairs)
  const int unit0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bar_empty = bar_full + Cfg::STAGES;
  uint64_t* bar_tfull = bar_empty + Cfg::STAGES;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* keep_stage = smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256;      // EPI_MFB: keep tiles; else: output staging

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = p.full_units + (p.batch * p.m_blocks * p.n_blocks - p.full_units) * p.k_split;   // units

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&bar_full[s], CTA2 ? 2 : 1);          // pair: leader's expect_tx arrive + the peer's remote arrive
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], CTA2 ? 16 : 8);       // epilogue warps of both CTAs release the leader's MMA thread
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (one lane) ===============================
    if (lane == 0) {
      tma_prefetch_desc(&tma_a);
      tma_prefetch_desc(&tma_b);
      int s = 0;
      uint32_t ph = 0;
      constexpr int TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
      VQA_DBG(long long prod_wait = 0; const long long prod_t0 = clock64();)
      for (int t = unit0; t < total_tiles; t += grid_units) {
        const Unit un = unit_decode(p, t);
        const int n_blk = un.n_blk, m_blk = un.m_blk, bz = un.bz, kb0 = un.kb0, kb1 = un.kb1;
        const int m_row0 = m_blk * TILE_M + (int)cta_rank * BLOCK_M;          // this CTA's 128 rows of A
        const int n_row0 = n_blk * BN + (int)cta_rank * Cfg::BN_CTA;          // this CTA's share of the B tile
        for (int kb = kb0; kb < kb1; ++kb) {
          VQA_DBG(const long long tq0 = clock64();)
          mbar_wait(&bar_empty[s], ph ^ 1);
          VQA_DBG(prod_wait += clock64() - tq0;)
          if constexpr (CTA2) {
            if (cta_rank == 0) mbar_expect_tx(&bar_full[s], 2 * Cfg::STAGE_BYTES);   // bytes of both CTAs
            else mbar_arrive_cluster(&bar_full[s], 0);
          } else {
            mbar_expect_tx(&bar_full[s], Cfg::STAGE_BYTES);
          }
          uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          auto load = [&](void* dst, const CUtensorMap* map, int c0, int c1) {
            if constexpr (CTA2) tma_load_3d_2cta(dst, map, &bar_full[s], c0, c1, bz);
            else tma_load_3d(dst, map, &bar_full[s], c0, c1, bz);
          };
          if (p.a_mn) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) load(sa + i * 8192, &tma_a, m_row0 + i * 64, kb * BLOCK_K);
          } else {
            load(sa, &tma_a, kb * BLOCK_K, m_row0);
          }
          if (p.b_mn) {
            if constexpr (Cfg::BN_CTA % 64 == 0) {
#pragma unroll
              for (int i = 0; i < Cfg::BN_CTA / 64; ++i) load(sb + i * 8192, &tma_b, n_row0 + i * 64, kb * BLOCK_K);
            }
          } else {
            load(sb, &tma_b, kb * BLOCK_K, n_row0);
          }
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
      }
      VQA_DBG(if (p.dbg != nullptr && blockIdx.x < 2) {
        p.dbg[blockIdx.x * 8 + 0] = (unsigned long long)prod_wait;
        p.dbg[blockIdx.x * 8 + 1] = (unsigned long long)(clock64() - prod_t0);
      })
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (one lane) ===============================
    if (lane == 0 && cta_rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      VQA_DBG(long long w_full = 0, w_tempty = 0; const long long mma_t0 = clock64();)
      for (int t = unit0; t < total_tiles; t += grid_units, ++it) {
        const Unit un = unit_decode(p, t);
        const int kb0 = un.kb0, kb1 = un.kb1;
        const int as = it & 1;
        const uint32_t aph = (it >> 1) & 1;
        VQA_DBG(const long long tq1 = clock64();)
        mbar_wait(&bar_tempty[as], aph ^ 1);          // epilogue has drained this accumulator
        VQA_DBG(w_tempty += clock64() - tq1;)
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * Cfg::ACC_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          VQA_DBG(const long long tq2 = clock64();)
          mbar_wait(&bar_full[s], ph);                // TMA bytes have landed
          VQA_DBG(w_full += clock64() - tq2;)
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::STAGE_BYTES) >> 4;
          const uint32_t b_addr = a_addr + (Cfg::A_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = p.a_desc_hi | (uint64_t)((a_addr + k * p.a_kadv) & 0x3FFF);
            const uint64_t db = p.b_desc_hi | (uint64_t)((b_addr + k * p.b_kadv) & 0x3FFF);
            if constexpr (CTA2) umma_bf16_2cta(tmem_d, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, da, db, p.idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (CTA2) umma_commit_2cta(&bar_empty[s]);
          else umma_commit(&bar_empty[s]);
          if (++s == Cfg::STAGES) { s = 0; ph ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CTA2) umma_commit_2cta(&bar_tfull[as]);
        else umma_commit(&bar_tfull[as]);
      }
      VQA_DBG(if (p.dbg != nullptr && blockIdx.x == 0) {
        p.dbg[2] = (unsigned long long)w_full;
        p.dbg[3] = (unsigned long long)w_tempty;
        p.dbg[4] = (unsigned long long)(clock64() - mma_t0);
      })
    }
  } else {
    // =============================== epilogue warps ===============================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;                 // which of the quadrant's two warps
    const int row_in_tile = quad * 32 + lane;
    const uint32_t st_seed = effective_seed(p.st_drop_seed, p.seed_dev);
    const uint32_t mfb_seed = effective_seed(p.drop_seed, p.seed_dev);
    int it = 0;
    constexpr int TILE_M = CTA2 ? 2 * BLOCK_M : BLOCK_M;
    for (int t = unit0; t < total_tiles; t += grid_units, ++it) {
      const Unit un = unit_decode(p, t);
      const int n_blk = un.n_blk, m_blk = un.m_blk, bz = un.bz;
      const int as = it & 1;
      const uint32_t aph = (it >> 1) & 1;
      const int m = m_blk * TILE_M + (int)cta_rank * BLOCK_M + row_in_tile;
      const bool row_ok = m < p.M;
      const int n0 = n_blk * BN;
      mbar_wait(&bar_tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * Cfg::ACC_STRIDE;

      if constexpr (EPI == EPI_STORE || EPI == EPI_ATOMIC) {
        const int grp = row_ok ? (m / p.rows_per_group) : 0;
        const float rs = (p.row_scale != nullptr && row_ok) ? p.row_scale[grp] : 1.0f;
        float dot_acc = 0.f;
        uint8_t* my_out = keep_stage + (warp - 2) * (32 * 128);               // this warp's 4 KB output staging tile
        const int m_base = m_blk * TILE_M + (int)cta_rank * BLOCK_M + quad * 32;   // first row of the warp's 32 rows
#pragma unroll 1
        for (int c0 = half * 32; c0 < BN; c0 += 64) {
          if (n0 + c0 >= p.N) break;                  // warp-uniform
          float v[32];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld_wait();
          const int n = n0 + c0;
          if constexpr (EPI == EPI_ATOMIC) {
            if (p.vec_ok && n + 32 <= p.N) {
              // 128-bit vector reductions (REDG.E.ADD.F32x4), one instruction = four whole 128-byte row segments
              stage_rows_f32(my_out, lane, v);
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3), jj = lane & 7;
                const float4 u = *reinterpret_cast<const float4*>(my_out + rr * 128 + ((jj ^ (rr & 7)) << 4));
                if (m_base + rr < p.M) {
                  float* dst = reinterpret_cast<float*>(p.C) + (long long)bz * p.c_bstride +
                               (long long)(m_base + rr) * p.ldc + n + jj * 4;
                  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(u.x), "f"(u.y), "f"(u.z),
                               "f"(u.w)
                               : "memory");
                }
              }
              __syncwarp();
            } else if (row_ok) {
              float* crow = reinterpret_cast<float*>(p.C) + (long long)bz * p.c_bstride + (long long)m * p.ldc + n;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (n + i < p.N) atomicAdd(crow + i, v[i]);
            }
          } else {
            const long long boff = (long long)bz * p.c_bstride + (long long)m * p.ldc + n;
            if (p.add == nullptr && p.act <= 1) {
              // fast path (every Linear / 1x1 conv of the MFB / MFH nets; hieCoAtten's img_emb): out = relu?(acc * rs +
              // bias), then the optional always-on dropout of hieCoAtten.py:26
              if (p.bias != nullptr) {
                if (n + 32 <= p.N) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
                    v[4 * q + 0] = fmaf(v[4 * q + 0], rs, b4.x);
                    v[4 * q + 1] = fmaf(v[4 * q + 1], rs, b4.y);
                    v[4 * q + 2] = fmaf(v[4 * q + 2], rs, b4.z);
                    v[4 * q + 3] = fmaf(v[4 * q + 3], rs, b4.w);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i) v[i] = (n + i < p.N) ? fmaf(v[i], rs, __ldg(p.bias + n + i)) : 0.f;
                }
              } else if (p.row_scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= rs;
              }
              if (p.act == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
              if (p.st_drop_thresh16 != 0) {
                const uint32_t grow = (uint32_t)(bz * p.M + m);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const uint32_t rb = dropout_bits(st_seed, grow, (uint32_t)((n + i) >> 1));
                  v[i] = ((rb & 0xFFFFu) >= p.st_drop_thresh16) ? v[i] * p.st_drop_scale : 0.f;
                  v[i + 1] = ((rb >> 16) >= p.st_drop_thresh16) ? v[i + 1] * p.st_drop_scale : 0.f;
                }
              }
            } else {
              // general path (hieCoAtten's per-sample products: residual addend, tanh, dropout): same steps, the bias
              // and the addend fetched as 128-bit words whenever the chunk is whole
              const bool whole = (n + 32 <= p.N);
              if (p.row_scale != nullptr) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] *= rs;
              }
              if (p.bias != nullptr) {
                if (whole) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
                    v[4 * q + 0] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (n + i < p.N) v[i] += __ldg(p.bias + n + i);
                }
              }
              if (p.add != nullptr && row_ok) {
                if (whole && p.vec_ok && !p.add_bf16) {
                  const float4* a4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.add) + boff);
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const float4 t4 = __ldg(a4 + q);
                    v[4 * q + 0] += t4.x; v[4 * q + 1] += t4.y; v[4 * q + 2] += t4.z; v[4 * q + 3] += t4.w;
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 32; ++i)
                    if (n + i < p.N)
                      v[i] += p.add_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.add)[boff + i])
                                         : reinterpret_cast<const float*>(p.add)[boff + i];
                }
              }
              if (p.act == 1) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              } else if (p.act == 2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
              }
              if (p.st_drop_thresh16 != 0) {
                const uint32_t grow = (uint32_t)(bz * p.M + m);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const uint32_t rb = dropout_bits(st_seed, grow, (uint32_t)((n + i) >> 1));
                  v[i] = ((rb & 0xFFFFu) >= p.st_drop_thresh16) ? v[i] * p.st_drop_scale : 0.f;
                  v[i + 1] = ((rb >> 16) >= p.st_drop_thresh16) ? v[i + 1] * p.st_drop_scale : 0.f;
                }
              }
            }
            const bool full = (n + 32 <= p.N);
            if (row_ok && p.dot_with != nullptr) {
              const __nv_bfloat16* drow = p.dot_with + (long long)bz * p.c_bstride + (long long)m * p.ld_dot + n;
              if (full && p.vec_ok) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = __ldg(reinterpret_cast<const uint4*>(drow) + q);
                  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    dot_acc += v[q * 8 + 2 * j] * bf16_lo(w[j]);
                    dot_acc += v[q * 8 + 2 * j + 1] * bf16_hi(w[j]);
                  }
                }
              } else {
                for (int i = 0; i < 32; ++i)
                  if (n + i < p.N) dot_acc += v[i] * __bfloat162float(drow[i]);
              }
            }
            if (full && p.vec_ok) {
              // coalesced write-out through the warp's staging tile (see GemmCfg::OUT_STAGE_BYTES)
              if (p.c_bf16) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 u;
                  u.x = pack_bf16(v[q * 8 + 0], v[q * 8 + 1]);
                  u.y = pack_bf16(v[q * 8 + 2], v[q * 8 + 3]);
                  u.z = pack_bf16(v[q * 8 + 4], v[q * 8 + 5]);
                  u.w = pack_bf16(v[q * 8 + 6], v[q * 8 + 7]);
                  *reinterpret_cast<uint4*>(my_out + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = u;
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int rr = i * 8 + (lane >> 2), jj = lane & 3;
                  const uint4 u = *reinterpret_cast<const uint4*>(my_out + rr * 64 + ((jj ^ ((rr >> 1) & 3)) << 4));
                  if (m_base + rr < p.M)
                    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)bz * p.c_bstride +
                                              (long long)(m_base + rr) * p.ldc + n + jj * 8) = u;
                }
              } else {
                stage_rows_f32(my_out, lane, v);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const int rr = i * 4 + (lane >> 3), jj = lane & 7;
                  const float4 u = *reinterpret_cast<const float4*>(my_out + rr * 128 + ((jj ^ (rr & 7)) << 4));
                  if (m_base + rr < p.M)
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (long long)bz * p.c_bstride +
                                               (long long)(m_base + rr) * p.ldc + n + jj * 4) = u;
                }
              }
              __syncwarp();
            } else if (row_ok) {
              if (p.c_bf16) {
                __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + boff;
                for (int i = 0; i < 32; ++i)
                  if (n + i < p.N) crow[i] = __float2bfloat16_rn(v[i]);
              } else {
                float* crow = reinterpret_cast<float*>(p.C) + boff;
                for (int i = 0; i < 32; ++i)
                  if (n + i < p.N) crow[i] = v[i];
              }
            }
          }
        }
        if constexpr (EPI == EPI_STORE) {
          if (p.dot_out != nullptr) {
            // rows of a warp usually belong to one group (sample): one atomic per warp then
            const int g0 = __shfl_sync(0xffffffffu, grp, 0);
            const bool uni = __all_sync(0xffffffffu, grp == g0);
            if (uni) {
              const float s = warp_sum(row_ok ? dot_acc : 0.f);
              if (lane == 0) atomicAdd(p.dot_out + g0, s);
            } else if (row_ok) {
              atomicAdd(p.dot_out + grp, dot_acc);
            }
          }
        }
      } else {
        // ---------------- EPI_MFB: Hadamard with Q, (dropout), sum over k=5, signed sqrt, sum|z| ----------------
        static_assert(EPI != EPI_MFB || BN % 80 == 0, "MFB epilogue: two warps x 40-column chunks (8 groups of k=5)");
        const int grp = row_ok ? (m / p.rows_per_group) : 0;
        const float* qrow = p.mfb_q + (long long)grp * p.mfb_ldq;
        const float* erow = (MFBX && p.mfb_extra != nullptr) ? p.mfb_extra + (long long)grp * p.mfb_ldq : nullptr;
        // bf16 keep tiles go through smem so that each row leaves the SM as 80 contiguous bytes
        const bool stage_keep = (p.mfb_keep != nullptr) && !p.mfb_keep_f32 && (p.N % 8 == 0);
        uint8_t* my_stage = keep_stage + (warp - 2) * (32 * Cfg::KEEP_PITCH);
        float abs_acc = 0.f;
        const int nseg = MFBX ? p.N / p.mfb_seg_cols : 1;
        int cur_seg = MFBX ? -1 : 0;
        // sum |z| of the rows of one norm segment: one atomic per warp when its rows share a group (sample)
        auto flush_ssq = [&](int seg, float acc) {
          const int g0 = __shfl_sync(0xffffffffu, grp, 0);
          const bool uni = __all_sync(0xffffffffu, grp == g0);
          if (uni) {
            const float sres = warp_sum(row_ok ? acc : 0.f);
            if (lane == 0) atomicAdd(p.mfb_ssq + (long long)g0 * nseg + seg, sres);
          } else if (row_ok) {
            atomicAdd(p.mfb_ssq + (long long)grp * nseg + seg, acc);
          }
        };
#pragma unroll 1
        for (int c0 = half * 40; c0 < BN; c0 += 80) {
          if (n0 + c0 >= p.N) break;                  // warp-uniform
          float v[40];
          tmem_ld16(taddr + c0, v);
          tmem_ld16(taddr + c0 + 16, v + 16);
          tmem_ld8(taddr + c0 + 32, v + 32);
          tmem_ld_wait();
          const int n = n0 + c0;                      // multiple of 40 -> 16-byte aligned float4 loads
          if constexpr (MFBX) {
            const int seg = n / p.mfb_seg_cols;       // warp-uniform; a 40-column chunk never straddles a segment
            if (seg != cur_seg) {
              if (cur_seg >= 0) flush_ssq(cur_seg, abs_acc);
              abs_acc = 0.f;
              cur_seg = seg;
            }
          }
          if (row_ok) {
            float z[8];
#pragma unroll
            for (int q = 0; q < 10; ++q) {            // 10 float4 = 40 columns
              if (n + q * 4 < p.N) {                  // N % 20 == 0 -> whole float4 in range
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + q);
                float4 q4 = __ldg(reinterpret_cast<const float4*>(qrow + n) + q);
                if constexpr (MFBX) {
                  if (erow != nullptr) {
                    const float4 e4 = __ldg(reinterpret_cast<const float4*>(erow + n) + q);
                    q4.x *= e4.x; q4.y *= e4.y; q4.z *= e4.z; q4.w *= e4.w;
                  }
                }
                v[q * 4 + 0] += b4.x; v[q * 4 + 1] += b4.y; v[q * 4 + 2] += b4.z; v[q * 4 + 3] += b4.w;
                if (p.drop_thresh16 != 0) {
                  const uint32_t r0 = dropout_bits(mfb_seed, (uint32_t)m, (uint32_t)((n + q * 4) >> 1));
                  const uint32_t r1 = dropout_bits(mfb_seed, (uint32_t)m, (uint32_t)((n + q * 4) >> 1) + 1);
                  v[q * 4 + 0] = ((r0 & 0xFFFFu) >= p.drop_thresh16) ? v[q * 4 + 0] * p.drop_scale : 0.f;
                  v[q * 4 + 1] = ((r0 >> 16) >= p.drop_thresh16) ? v[q * 4 + 1] * p.drop_scale : 0.f;
                  v[q * 4 + 2] = ((r1 & 0xFFFFu) >= p.drop_thresh16) ? v[q * 4 + 2] * p.drop_scale : 0.f;
                  v[q * 4 + 3] = ((r1 >> 16) >= p.drop_thresh16) ? v[q * 4 + 3] * p.drop_scale : 0.f;
                }
                if (p.mfb_keep != nullptr) {
                  if (p.mfb_keep_f32) {
                    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.mfb_keep) + (long long)m * p.N + n + q * 4) =
                        make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                  } else {
                    uint2 u;
                    u.x = pack_bf16(v[q * 4 + 0], v[q * 4 + 1]);
                    u.y = pack_bf16(v[q * 4 + 2], v[q * 4 + 3]);
                    if (stage_keep)
                      *reinterpret_cast<uint2*>(my_stage + lane * Cfg::KEEP_PITCH + q * 8) = u;
                    else
                      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.mfb_keep) + (long long)m * p.N + n + q * 4) = u;
                  }
                }
                v[q * 4 + 0] *= q4.x; v[q * 4 + 1] *= q4.y; v[q * 4 + 2] *= q4.z; v[q * 4 + 3] *= q4.w;
                if constexpr (MFBX) {
                  if (p.mfb_prod != nullptr)
                    *reinterpret_cast<float4*>(p.mfb_prod + (long long)m * p.N + n + q * 4) =
                        make_float4(v[q * 4 + 0], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                }
              } else {
                v[q * 4 + 0] = 0.f; v[q * 4 + 1] = 0.f; v[q * 4 + 2] = 0.f; v[q * 4 + 3] = 0.f;
              }
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float zz = (v[5 * g] + v[5 * g + 1]) + (v[5 * g + 2] + v[5 * g + 3]) + v[5 * g + 4];
              abs_acc += fabsf(zz);
              z[g] = copysignf(sqrtf(fabsf(zz)), zz);
            }
            const int o0 = n / 5;                     // multiple of 8
            const int No = p.N / 5;
            if (p.mfb_y_bf16) {
              __nv_bfloat16* yrow = reinterpret_cast<__nv_bfloat16*>(p.mfb_y) + (long long)m * p.mfb_ldy + o0;
              if (o0 + 8 <= No && p.vec_ok) {
                uint4 u;
                u.x = pack_bf16(z[0], z[1]); u.y = pack_bf16(z[2], z[3]);
                u.z = pack_bf16(z[4], z[5]); u.w = pack_bf16(z[6], z[7]);
                *reinterpret_cast<uint4*>(yrow) = u;
              } else {
                for (int g = 0; g < 8; ++g)
                  if (o0 + g < No) yrow[g] = __float2bfloat16_rn(z[g]);
              }
            } else {
              float* yrow = reinterpret_cast<float*>(p.mfb_y) + (long long)m * p.mfb_ldy + o0;
              if (o0 + 8 <= No && p.vec_ok) {
                reinterpret_cast<float4*>(yrow)[0] = make_float4(z[0], z[1], z[2], z[3]);
                reinterpret_cast<float4*>(yrow)[1] = make_float4(z[4], z[5], z[6], z[7]);
              } else {
                for (int g = 0; g < 8; ++g)
                  if (o0 + g < No) yrow[g] = z[g];
              }
            }
          }
          if (stage_keep) {
            __syncwarp();
            const int m_base = m_blk * TILE_M + (int)cta_rank * BLOCK_M + quad * 32;
            __nv_bfloat16* kbase = reinterpret_cast<__nv_bfloat16*>(p.mfb_keep);
#pragma unroll
            for (int i = 0; i < 5; ++i) {
              const int qi = lane + 32 * i;           // 160 16-byte pieces: 32 rows x 5
              const int rr = qi / 5, cc = qi % 5;
              const uint4 u = *reinterpret_cast<const uint4*>(my_stage + rr * Cfg::KEEP_PITCH + cc * 16);
              if (m_base + rr < p.M && n + cc * 8 < p.N)
                *reinterpret_cast<uint4*>(kbase + (long long)(m_base + rr) * p.N + n + cc * 8) = u;
            }
            __syncwarp();
          }
        }
        if (cur_seg >= 0) flush_ssq(cur_seg, abs_acc);
      }
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTA2) mbar_arrive_cluster(&bar_tempty[as], 0);
        else mbar_arrive(&bar_tempty[as]);
      }
    }
  }

  tc_fence_before();
  __syncwarp();                                   // re-converge the single-lane role warps before an .aligned barrier
  if constexpr (CTA2) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CTA2) tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vqa
