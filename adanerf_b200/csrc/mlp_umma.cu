// tcgen05 / TMEM fused MLP kernel (see mlp_umma.cuh for the design).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <type_traits>

#include "mlp_umma.cuh"
#include "posenc.cuh"
#include "ptx.cuh"

namespace adn {

template <int NSPLIT>
struct MlpCfg {
  static constexpr int kNB = (NSPLIT == 2) ? 4 : 5;            // activation blocks per slot and term
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// ReLU folded into the fp32 -> bf16x2 conversion (one instruction for two elements); a -> low half.
__device__ __forceinline__ uint32_t pack_relu_bf16x2(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}

// Epilogue kinds (derived from the layer flags once per layer, so the per-element code is branch free).
enum : int { EK_ACT_RELU = 0, EK_ACT_RELU_ALPHA = 1, EK_ACT_LINEAR = 2, EK_FINAL_RGB = 3, EK_FINAL_RAW = 4, EK_FINAL_RAW_STAGED = 5 };

__device__ __forceinline__ int epilogue_kind(uint8_t flags) {
  if (flags & LF_FINAL_RAW) return EK_FINAL_RAW;
  if (flags & LF_FINAL_RGB) return EK_FINAL_RGB;
  if (!(flags & LF_RELU)) return EK_ACT_LINEAR;
  return (flags & LF_ALPHA_DOT) ? EK_ACT_RELU_ALPHA : EK_ACT_RELU;
}

// Epilogue for 32 consecutive accumulator columns [c, c+32) of one row (thread = row).
template <int NSPLIT, int KIND>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], int c, const MlpLayer& L, const MlpProgram& prog,
                                               uint32_t act_hi, uint32_t act_lo, int row_in_tile, long long grow,
                                               long long rows, float* __restrict__ out, float& alpha, float (&rgb)[3]) {
  constexpr bool kRelu = (KIND == EK_ACT_RELU || KIND == EK_ACT_RELU_ALPHA || KIND == EK_FINAL_RGB);
  constexpr bool kAct = (KIND == EK_ACT_RELU || KIND == EK_ACT_RELU_ALPHA || KIND == EK_ACT_LINEAR);
  float v[32];
  if (NSPLIT == 1 && kAct) {
    // plain-bf16 nets: hidden-layer biases are stored as bf16 pairs (bias_off = word offset of the layer's packed
    // vector).  Register-indexed constant loads are 64-bit at most and the constant path, not the ALU, limits this
    // epilogue (dropping the loads altogether: shading kernel 5.04 -> 4.54 ms): 8 loads + 32 shifts / masks replace 16
    // loads.  bf16 rounding of a bias (|b| < 0.1: < 2e-4 absolute) is far below the bf16 rounding of the activations.
    const uint2* b2 = reinterpret_cast<const uint2*>(prog.side) + ((L.bias_off + (c >> 1)) >> 1);   // warp-uniform
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint2 w = b2[j];
      v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + __uint_as_float(w.x << 16);
      v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + __uint_as_float(w.x & 0xFFFF0000u);
      v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + __uint_as_float(w.y << 16);
      v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + __uint_as_float(w.y & 0xFFFF0000u);
    }
  } else {
    const float4* bias = reinterpret_cast<const float4*>(prog.side) + ((L.bias_off + c) >> 2);   // warp-uniform, constant bank
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = bias[j];
      v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
    }
  }
  constexpr bool kFusedReluPack = (NSPLIT == 1 && KIND == EK_ACT_RELU);   // relu inside the bf16 pack
  if (kRelu && !kFusedReluPack) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (KIND == EK_ACT_RELU_ALPHA) {
    // four independent partial sums: a single accumulator would be a 32-deep dependent FMA chain
    const float4* w4 = reinterpret_cast<const float4*>(prog.side) + ((prog.alpha_w_off + c) >> 2);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = w4[j];
      a0 = fmaf(v[4 * j + 0], w.x, a0);
      a1 = fmaf(v[4 * j + 1], w.y, a1);
      a2 = fmaf(v[4 * j + 2], w.z, a2);
      a3 = fmaf(v[4 * j + 3], w.w, a3);
    }
    alpha += (a0 + a1) + (a2 + a3);
  }
  if (kAct) {
    // columns [c, c+32) -> block out_blk0 + c/64, 16-byte chunks (c%64)/8 .. +3, XOR-swizzled by row%8
    const uint32_t blk = L.out_blk0 + (c >> 6);
    const uint32_t rbase = blk * kBlkBytes + (row_in_tile >> 3) * 1024u + (row_in_tile & 7) * 128u;
    const uint32_t cc0 = (c & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 hi;
      if (kFusedReluPack) {
        hi.x = pack_relu_bf16x2(v[8 * q + 0], v[8 * q + 1]);
        hi.y = pack_relu_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        hi.z = pack_relu_bf16x2(v[8 * q + 4], v[8 * q + 5]);
        hi.w = pack_relu_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      } else {
        hi.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
        hi.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
        hi.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
        hi.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
      }
      const uint32_t off = rbase + (((cc0 + q) ^ (row_in_tile & 7)) << 4);
      st_shared_v4(act_hi + off, hi.x, hi.y, hi.z, hi.w);
      if (NSPLIT == 2) {
        float l[8];
        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          l[2 * e + 0] = v[8 * q + 2 * e + 0] - __uint_as_float(hw[e] << 16);
          l[2 * e + 1] = v[8 * q + 2 * e + 1] - __uint_as_float(hw[e] & 0xFFFF0000u);
        }
        uint4 lo;
        lo.x = pack_bf16x2(l[0], l[1]);
        lo.y = pack_bf16x2(l[2], l[3]);
        lo.z = pack_bf16x2(l[4], l[5]);
        lo.w = pack_bf16x2(l[6], l[7]);
        st_shared_v4(act_lo + off, lo.x, lo.y, lo.z, lo.w);
      }
    }
  }
  if (KIND == EK_FINAL_RAW_STAGED) {
    // fp32 row slice -> shared-memory staging (two 32 KB pieces that the next tile's input load does not touch);
    // 16-byte units are XOR-swizzled by the row so that both these row-wise writes and the later coalesced row
    // reads are bank-conflict free.  The block-wide copy to global happens after the layer (see kernel).
    const uint32_t base = act_hi + ((row_in_tile < 64) ? 2u : 6u) * kBlkBytes + uint32_t(row_in_tile & 63) * 512u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t unit = (uint32_t(c >> 2) + j) ^ uint32_t(row_in_tile & 31);
      st_shared_v4(base + unit * 16u, __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]), __float_as_uint(v[4 * j + 2]),
                   __float_as_uint(v[4 * j + 3]));
    }
  }
  if (KIND == EK_FINAL_RAW) {
    if (grow < rows) {
      float4* o4 = reinterpret_cast<float4*>(out + grow * prog.out_cols + c);
#pragma unroll
      for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  if (KIND == EK_FINAL_RGB) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4* w4 = reinterpret_cast<const float4*>(prog.side) + ((prog.rgb_w_off + k * 128 + c) >> 2);
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w = w4[j];
        a0 = fmaf(v[4 * j + 0], w.x, a0);
        a1 = fmaf(v[4 * j + 1], w.y, a1);
        a2 = fmaf(v[4 * j + 2], w.z, a2);
        a3 = fmaf(v[4 * j + 3], w.w, a3);
      }
      rgb[k] += (a0 + a1) + (a2 + a3);
    }
  }
}

// All accumulator columns [c0, c0 + span) of one layer for this thread's row.
template <int NSPLIT, int KIND>
__device__ __forceinline__ void epilogue_layer(uint32_t taddr, int c0, int span, const MlpLayer& L, const MlpProgram& prog,
                                               uint32_t act_hi, uint32_t act_lo, int row_in_tile, long long grow,
                                               long long rows, float* __restrict__ out, float& alpha, float (&rgb)[3]) {
#pragma unroll 1
  for (int c = c0; c < c0 + span; c += 32) {
    uint32_t ra[32];
    tmem_ld32(taddr + c, ra);
    tc_wait_ld();
    epilogue_chunk<NSPLIT, KIND>(ra, c, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
  }
}

// Ring geometry.  CG = 1: plain bf16 -> 2 stages of [256 x 64] (one K block, both N halves); split precision -> 3 stages
// of [128 x 64] hi + lo.  CG = 2 (CTA pair, cta_group::2 MMAs with M = 256): every CTA streams only ITS half of the
// B rows of each stage, so the same 64 / 96 KB hold twice as many stages (4 / 6): three or more weight stages are in
// flight while one is consumed, and each SM pulls half the weight bytes from L2.
// Ring stages a layer consumes.  Split precision: one per (K block, N half).  Plain bf16: one per K block for
// 256-wide layers; 128-wide layers pack TWO K blocks into a stage (same bytes as one wide K block), so every
// barrier round trip of the issuer still carries 8 x 64 = 512 tensor cycles of work.
template <int NSPLIT>
__device__ __forceinline__ int layer_stages(const MlpLayer& L) {
  if (NSPLIT == 2) return int(L.n_kb) * int(L.n_half);
  return (L.n_half == 2) ? int(L.n_kb) : (int(L.n_kb) + 1) / 2;
}

template <int NSPLIT, int CG>
struct RingCfg {
  static constexpr int kStageBytes = 2 * kBlkBytes / CG;
#ifndef ADN_BF16_PAIR_STAGES
#define ADN_BF16_PAIR_STAGES 4   // ring stages of the shading kernel's CTA pair (sensitivity experiments: 3)
#endif
  static constexpr int kStages = (NSPLIT == 1 && CG == 2) ? ADN_BF16_PAIR_STAGES : ((NSPLIT == 2) ? 3 : 2) * CG;
};

template <int NSPLIT, int NG, int CG>
constexpr size_t mlp_smem_layout_bytes() {
  return size_t(NG) * NSPLIT * MlpCfg<NSPLIT>::kNB * kBlkBytes + size_t(RingCfg<NSPLIT, CG>::kStages) * RingCfg<NSPLIT, CG>::kStageBytes +
         512 /*barriers*/ + 1024 /*alignment slack*/;
}

template <int NSPLIT, int NG, int CG, bool ENC = false>
__global__ void __launch_bounds__(ENC ? kMlpEncThreads : kMlpThreads, 1)
mlp_umma_kernel(const __grid_constant__ MlpProgram prog, const uint8_t* __restrict__ wblob,
                const uint8_t* __restrict__ in_tiles, float* __restrict__ out,
                const long long* __restrict__ rows_dev, long long rows_host, int* err_flag, long long* trace,
                const __grid_constant__ EncodeParams enc) {
  static_assert(!ENC || NSPLIT == 1, "fused input encoder: shading net (plain bf16) only");
  using Cfg = MlpCfg<NSPLIT>;
  using Ring = RingCfg<NSPLIT, CG>;
  constexpr int NB = Cfg::kNB;
  constexpr int STAGES = Ring::kStages;
  constexpr int STAGE_BYTES = Ring::kStageBytes;
  constexpr int HALF = kBlkBytes / CG;   // bytes of one [128 x 64] B tile held by one CTA
  constexpr int EW = 16;        // epilogue warps: every one of them serves all NG tile slots in turn
  constexpr int QW = EW / 4;    // warps sharing one TMEM lane quarter (they split the columns)
  constexpr int CW = 128 / QW;  // accumulator columns per warp and N half
  // ENC: warp 16 encodes the tile inputs.  (As warp 0 -- lowest issue priority -- it starves behind its sub-partition's
  // epilogue warps and the MMAs end up waiting for input images: measured 6.3 ms against 5.8 ms.)
  constexpr int kEncWarp = 16;
  constexpr int kProducerWarp = ENC ? 17 : 16, kHelperWarp = kProducerWarp + 1, kMmaWarp = kProducerWarp + 2;   // highest warp ids: favoured by the issue arbiter
  constexpr int kBarSlot = 3, kBarStage = 5;   // named barrier ids (1, 2 are used by the epilogue warps)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;                                                   // [NG][NSPLIT][NB] blocks
  uint8_t* ring = act + size_t(NG) * NSPLIT * NB * kBlkBytes;            // [STAGES] stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + size_t(STAGES) * STAGE_BYTES);
  uint64_t* w_full = bars;                     // [STAGES] this CTA's share of the stage has landed
  uint64_t* w_empty = w_full + STAGES;         // [STAGES] the MMAs reading the stage have retired (both CTAs)
  uint64_t* peer_full = w_empty + STAGES;      // [STAGES] leader only: the peer CTA's share has landed
  uint64_t* acc_full = peer_full + STAGES;     // [NG]  all MMAs of the layer have retired
  uint64_t* act_ready = acc_full + NG;         // [NG]  this CTA's epilogue warps are done (TMEM free, A operand written)
  uint64_t* peer_act = act_ready + NG;         // [NG]  leader only: the peer's act_ready, forwarded by its helper warp
  uint64_t* in_full = peer_act + NG;           // [NG]  this CTA's tile input has landed
  uint64_t* peer_in = in_full + NG;            // [NG]  leader only: the peer's tile input has landed
  uint64_t* enc_ready = peer_in + NG;          // [NG]  ENC: the encoder warp has written the next tile's input image
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(enc_ready + NG);
  volatile int* tiles_done = reinterpret_cast<volatile int*>(tmem_slot + 2);   // [NG] ENC: tiles of the slot whose last layer has retired

  // warp index through a lane-0 broadcast: tells the compiler it is warp uniform, so the role branches
  // (and everything indexed by loop counters inside them) stay on the uniform datapath
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const long long n_units = gridDim.x / CG;   // persistent work units: CTAs or CTA pairs
  const long long unit = blockIdx.x / CG;

  const long long rows = rows_dev ? *rows_dev : rows_host;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
      mbar_init(&peer_full[s], 1);
    }
    for (int g = 0; g < NG; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&act_ready[g], EW);
      mbar_init(&peer_act[g], 1);
      mbar_init(&in_full[g], 1);
      mbar_init(&peer_in[g], 1);
      mbar_init(&enc_ready[g], 1);
      tiles_done[g] = 0;
    }
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc_cg<CG>(tmem_slot, 512);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto act_ptr = [&](int g, int term, int blk) -> uint8_t* {
    return act + (size_t(g * NSPLIT + term) * NB + blk) * kBlkBytes;
  };
  // ENC: this CTA's input images in global memory (L2 resident), double buffered per slot: [P block | V block]
  auto enc_scratch = [&](int g, long long iter) -> uint8_t* {
    return enc.scratch + ((size_t(blockIdx.x) * NG + g) * 2 + size_t(iter & 1)) * (2 * kBlkBytes);
  };
  // first tile of the unit's tile group for (iter, slot); CTA `cta_rank` of a pair owns tile first + cta_rank
  auto first_tile = [&](long long iter, int g) -> long long { return ((iter * n_units + unit) * NG + g) * CG; };
  // Optional timeline (debug): a few warps of CTA 0 record (clock, code) pairs into private regions of `trace`
  // (region r: words [r*8192, (r+1)*8192), word 0 = count); code = slot<<16 | layer<<8 | event.  Off (nullptr)
  // in normal operation.
  const bool tracing = (trace != nullptr) && (blockIdx.x == 0);
  int tr_n = 0;
  auto tr = [&](int region, int g, int l, int ev) {
    if (tracing && tr_n < 4000) {
      long long* base = trace + region * 8192;
      base[2 + 2 * tr_n] = clock64();
      base[3 + 2 * tr_n] = (long long)((g << 16) | (l << 8) | ev);
      base[0] = ++tr_n;
    }
  };

  if (warp == kProducerWarp) {
    // ===================================================================== weight producer
    // Streams this CTA's share of every weight stage: stage i of a layer is [n_half x 128 rows x 64] (plain bf16)
    // or [128 rows x 64] hi + lo (split); with CG = 2 each CTA takes rows [rank*64, +64) of every 128-row tile.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter, 0) >= n_tiles) break;
        for (int l = 0; l < prog.n_layers; ++l) {
          const MlpLayer& L = prog.layers[l];
          const int n_st = layer_stages<NSPLIT>(L);
          const uint32_t src_stride = (NSPLIT == 2 || L.n_half == 2) ? uint32_t(2 * kBlkBytes) : uint32_t(kBlkBytes);
          for (int g = 0; g < NG; ++g) {
            if (first_tile(iter, g) >= n_tiles) continue;
            const uint8_t* src = wblob + size_t((blockIdx.x / CG) % prog.w_copies) * prog.w_stride + L.w_off;
            for (int i = 0; i < n_st; ++i) {
              if (l == 2) tr(5, g, stage, 8);
              mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 1);
              if (l == 2) tr(5, g, stage, 9);
              uint8_t* dst = ring + size_t(stage) * STAGE_BYTES;
              const uint8_t* s0 = src + size_t(i) * src_stride;
              if (NSPLIT == 2) {
                // [hi 128 rows | lo 128 rows] -> this CTA's row range of each
                mbar_arrive_expect_tx(&w_full[stage], 2 * HALF);
                bulk_g2s(dst, s0 + cta_rank * HALF, HALF, &w_full[stage]);
                bulk_g2s(dst + HALF, s0 + kBlkBytes + cta_rank * HALF, HALF, &w_full[stage]);
              } else if (L.n_half == 2) {
                if (CG == 2) {
                  // N = 256 over the pair: CTA r holds B rows [128 r, 128 r + 128) = N half r
                  mbar_arrive_expect_tx(&w_full[stage], kBlkBytes);
                  bulk_g2s(dst, s0 + cta_rank * kBlkBytes, kBlkBytes, &w_full[stage]);
                } else {
                  mbar_arrive_expect_tx(&w_full[stage], 2 * kBlkBytes);
                  bulk_g2s(dst, s0, 2 * kBlkBytes, &w_full[stage]);
                }
              } else {
                // 128-wide layer: K blocks 2i and 2i+1 (if present) share the stage, each as this CTA's row range
                const int nkb = (2 * i + 1 < int(L.n_kb)) ? 2 : 1;
                const uint8_t* s1 = src + size_t(2 * i) * kBlkBytes;
                mbar_arrive_expect_tx(&w_full[stage], uint32_t(nkb) * HALF);
                bulk_g2s(dst, s1 + cta_rank * HALF, HALF, &w_full[stage]);
                if (nkb == 2) bulk_g2s(dst + HALF, s1 + kBlkBytes + cta_rank * HALF, HALF, &w_full[stage]);
              }
              if (++stage == STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == kHelperWarp) {
    // =================================================================== barrier helper
    // Walks the same schedule as the MMA issuer one step ahead and turns every mbarrier round trip
    // (try_wait is ~100-200 cycles even when the phase has already completed) into a named-barrier
    // arrival, which the issuer consumes with a ~tens-of-cycles bar.sync.  In a CTA pair the peer's helper
    // forwards "my share has landed" to the leader with remote mbarrier arrivals.
    int stage = 0;
    uint32_t phase = 0;
    uint32_t in_phase[NG], ar_phase[NG];
#pragma unroll
    for (int g = 0; g < NG; ++g) in_phase[g] = ar_phase[g] = 0;
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
      for (int l = 0; l < prog.n_layers; ++l) {
        const MlpLayer& L = prog.layers[l];
        const int n_st = layer_stages<NSPLIT>(L);
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          if (first_tile(iter, g) >= n_tiles) continue;
          if (l == 0 || (L.flags & LF_WAIT_IN)) {
            mbar_wait(&in_full[g], in_phase[g], err_flag, 2);
            if (CG == 2) {
              if (leader) mbar_wait(&peer_in[g], in_phase[g], err_flag, 7);
              else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_in[g]), 0));
            }
            in_phase[g] ^= 1;
          }
          mbar_wait(&act_ready[g], ar_phase[g], err_flag, 3);
          if (CG == 2) {
            if (leader) mbar_wait(&peer_act[g], ar_phase[g], err_flag, 9);
            else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_act[g]), 0));
          }
          ar_phase[g] ^= 1;
          if (leader) named_bar_arrive(kBarSlot + g, 64);
          for (int i = 0; i < n_st; ++i) {
            if (lane == 0 && l == 2) tr(6, g, stage, 10);
            mbar_wait(&w_full[stage], phase, err_flag, 4);
            if (lane == 0 && l == 2) tr(6, g, stage, 11);
            if (CG == 2) {
              if (leader) mbar_wait(&peer_full[stage], phase, err_flag, 8);
              else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_full[stage]), 0));
              if (lane == 0 && l == 2) tr(6, g, stage, 12);
            }
            if (leader) named_bar_arrive(kBarStage + stage, 64);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ========================================================================== MMA issuer (leader CTA only)
    // The whole warp walks the schedule converged (every quantity below is warp uniform, so the
    // descriptors live in uniform registers); one elected lane issues the tcgen05 instructions.
    // Plain bf16: a 256-wide layer is issued as N = 256 MMAs (one stage = one K block).  CG = 2: M = 256 across the
    // CTA pair -- A rows and B rows of both CTAs are addressed by the same shared-memory offsets.
    if (leader) {
      constexpr uint32_t idesc128 = make_idesc_bf16(128 * CG, 128);
      constexpr uint32_t idesc256 = make_idesc_bf16(128 * CG, 256);
      int stage = 0;
      const uint64_t desc_hi = make_desc_sw128(0) & 0xFFFFFFFF00000000ull;   // constant upper word
      const uint32_t desc_lo_const = uint32_t(make_desc_sw128(0) & 0xFFFF0000ull);
      // low descriptor word of an address: a K step of 16 elements (+32 B) is "+2" on this word
      auto lo_of = [&](uint32_t addr) -> uint32_t { return desc_lo_const | (addr >> 4); };
      const uint32_t ring_lo = lo_of(smem_u32(ring));
      const uint32_t act_lo0 = lo_of(smem_u32(act));
      auto desc = [&](uint32_t lo) -> uint64_t { return desc_hi | uint64_t(lo); };
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter, 0) >= n_tiles) break;
        for (int l = 0; l < prog.n_layers; ++l) {
          const MlpLayer& L = prog.layers[l];
          const bool wide = (NSPLIT == 1) && (L.n_half == 2);
          // this layer's A block indices packed 4 bits each: the issue loop extracts them with a shift instead
          // of a dependent constant-bank load per K block
          uint32_t blks = 0;
#pragma unroll
          for (int kb = 0; kb < 6; ++kb) blks |= uint32_t(L.a_blk[kb] & 15) << (4 * kb);
#pragma unroll
          for (int g = 0; g < NG; ++g) {
            if (first_tile(iter, g) >= n_tiles) continue;
            if (lane == 0) tr(0, g, l, 0);
            named_bar_sync(kBarSlot + g, 64);   // helper has seen in_full / act_ready of this slot
            tc_fence_after();
            if (lane == 0) tr(0, g, l, 1);
            const uint32_t d0 = tmem_base + uint32_t(g * 256);
            for (int kb = 0; kb < L.n_kb; ++kb) {
              const uint32_t a_hi = act_lo0 + (uint32_t(g * NSPLIT * NB) + ((blks >> (4 * kb)) & 15u)) * (kBlkBytes >> 4);
              const uint32_t a_lo = a_hi + uint32_t((NSPLIT - 1) * NB) * (kBlkBytes >> 4);
              if (NSPLIT == 1) {
                if (lane == 0 && l == 2) tr(0, g, kb, 5);
                named_bar_sync(kBarStage + stage, 64);   // helper has seen w_full[stage] (of both CTAs)
                tc_fence_after();
                if (lane == 0 && l == 2) tr(0, g, kb, 6);
                const uint32_t b = ring_lo + uint32_t(stage) * (STAGE_BYTES >> 4);
                bool two = false;
                if (wide) {   // hot path: kept minimal, the issuer warp's instruction count is the kernel's clock
                  if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_bf16_cg<CG>(d0, desc(a_hi + 2 * k), desc(b + 2 * k), idesc256, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_cg<CG>(&w_empty[stage]);
                  }
                } else {
                  // 128-wide layer: this stage also carries the next K block (if the layer has one)
                  two = (kb + 1 < L.n_kb);
                  const uint32_t a_hi2 = act_lo0 + (uint32_t(g * NSPLIT * NB) + ((blks >> (4 * (kb + 1))) & 15u)) * (kBlkBytes >> 4);
                  if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                      umma_bf16_cg<CG>(d0, desc(a_hi + 2 * k), desc(b + 2 * k), idesc128, (kb > 0 || k > 0) ? 1u : 0u);
                    if (two) {
#pragma unroll
                      for (int k = 0; k < 4; ++k)
                        umma_bf16_cg<CG>(d0, desc(a_hi2 + 2 * k), desc(b + (HALF >> 4) + 2 * k), idesc128, 1u);
                    }
                    umma_commit_cg<CG>(&w_empty[stage]);
                  }
                }
                __syncwarp();
                if (lane == 0 && l == 2) tr(0, g, kb, 7);
                if (++stage == STAGES) stage = 0;
                if (two) ++kb;
              } else {
                for (int nh = 0; nh < L.n_half; ++nh) {
                  named_bar_sync(kBarStage + stage, 64);
                  tc_fence_after();
                  const uint32_t b_hi = ring_lo + uint32_t(stage) * (STAGE_BYTES >> 4);
                  const uint32_t b_lo = b_hi + (HALF >> 4);
                  const uint32_t d = d0 + uint32_t(nh * 128);
                  if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      umma_bf16_cg<CG>(d, desc(a_hi + 2 * k), desc(b_hi + 2 * k), idesc128, (kb > 0 || k > 0) ? 1u : 0u);
                      umma_bf16_cg<CG>(d, desc(a_lo + 2 * k), desc(b_hi + 2 * k), idesc128, 1u);
                      umma_bf16_cg<CG>(d, desc(a_hi + 2 * k), desc(b_lo + 2 * k), idesc128, 1u);
                    }
                    umma_commit_cg<CG>(&w_empty[stage]);
                  }
                  __syncwarp();
                  if (++stage == STAGES) stage = 0;
                }
              }
            }
            if (elect_one()) umma_commit_cg<CG>(&acc_full[g]);
            __syncwarp();
            if (lane == 0) tr(0, g, l, 2);
          }
        }
      }
    }
  } else if (ENC && warp == kEncWarp) {
    // ============================================================================ input encoder (stage 3, fused)
    // One warp, four tile rows per lane (rows lane + 32 q).  It runs up to two tiles ahead of the MMAs: for every tile
    // it computes the position block P and the view block V (RayMarchFromPoses.batch, features.py:458-479, same device
    // functions as stage3_kernel) and writes the swizzled bf16 image into this CTA's double-buffered scratch in global
    // memory (19 MB for the whole grid: L2 resident), from where the usual bulk copies fetch it -- P at the tile
    // start, V once the skip layer has retired.  The [M, 90] feature tensor never exists and the encoding is off the
    // critical path.  Buffer reuse: the image of tile t - 2 is dead when that tile's last layer has retired.
    auto store_row = [&](uint8_t* blk, int row, const float (&v)[3], bool live, auto L_tag) {
      constexpr int L = decltype(L_tag)::value;
      float f[64];
#pragma unroll
      for (int k = 0; k < 64; ++k) f[k] = 0.0f;
      if (live) posenc3<L>(v, f);
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float* q = f + ch * 8;
        const uint4 hi = make_uint4(bf16x2(q[0], q[1]), bf16x2(q[2], q[3]), bf16x2(q[4], q[5]), bf16x2(q[6], q[7]));
        *reinterpret_cast<uint4*>(blk + sw128_offset(uint32_t(row), uint32_t(ch * 8))) = hi;
      }
    };
    using L10 = std::integral_constant<int, 10>;
    using L4 = std::integral_constant<int, 4>;
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
#pragma unroll 1
      for (int g = 0; g < NG; ++g) {
        if (first_tile(iter, g) >= n_tiles) continue;
        const long long t = first_tile(iter, g) + cta_rank;
        if (iter >= 2) {   // the image buffer still belongs to tile iter - 2 until that tile's last layer has retired
          const long long t0 = clock64();
          while (tiles_done[g] < int(iter) - 1) {
            __nanosleep(200);
            if (clock64() - t0 > ADN_WATCHDOG_CYCLES) {
              if (err_flag) atomicExch(err_flag, 0x1000 + 10);
              asm volatile("trap;");
            }
          }
        }
        uint8_t* img = enc_scratch(g, iter);
#pragma unroll 1
        for (int q = 0; q < 4; ++q) {
          const int row = lane + 32 * q;
          const long long i = t * kTileM + row;
          const bool live = i < rows;
          float pos[3] = {0.f, 0.f, 0.f}, d[3] = {0.f, 0.f, 0.f};
          if (live) {
            long long ray;
            float zw;
            if (enc.ray_idx) {
              ray = enc.ray_idx[i];
              zw = enc.z[i];
            } else {
              ray = i / enc.K;
              zw = enc.zlut_dense[i - ray * enc.K];
            }
            // pos = o + d z, then normalization_inverse_sqrt_dist_centered (same operation order as stage3_kernel)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              d[a] = __ldg(enc.ray_d + 3 * ray + a);
              pos[a] = __fsub_rn(__fadd_rn(__ldg(enc.ray_o + 3 * ray + a), __fmul_rn(d[a], zw)), enc.c[a]);
            }
            const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(pos[0], pos[0]), __fmul_rn(pos[1], pos[1])), __fmul_rn(pos[2], pos[2])));
            const float den = __fmul_rn(enc.sqrt_max_depth, __fsqrt_rn(nrm));
#pragma unroll
            for (int a = 0; a < 3; ++a) pos[a] = __fdiv_rn(pos[a], den);
          }
          store_row(img, row, pos, live, L10{});                 // P: 63 position features + 0
          store_row(img + kBlkBytes, row, d, live, L4{});        // V: 27 direction features + zeros
        }
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive(&enc_ready[g]);
      }
    }
  } else {
    // ============================================================================ epilogue
    // All 16 warps serve BOTH tile slots, alternating (layer l slot 0, layer l slot 1, layer l+1 slot 0, ...): the
    // slots' accumulators become ready one after the other (the issuer alternates too), so each epilogue is done by
    // twice the warps in half the time and the slot's MMA -> epilogue -> MMA chain gets shorter than the two slots'
    // tensor time -- the tensor pipe, not the chain, then sets the pace.
    const int e = warp;                            // 0..15
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int sub = e >> 2;                        // which column slice of each N half this warp owns
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc_phase = 0;                        // bit g = parity of acc_full[g]
    float alpha_s0 = 0.0f, alpha_s1 = 0.0f;        // per-slot alpha partial (layer 7 -> final layer)
    // tile prologue of slot g: load its input tile and hand the (free) accumulator / activation buffers to layer 0
    auto tile_prologue = [&](long long iter, int g) {
      const long long t = first_tile(iter, g) + cta_rank;
      if (e == 0 && lane == 0) {
        if (ENC) {   // the encoder warp has written this tile's [P | V] image (generic proxy) -> read by the TMA engine
          mbar_wait(&enc_ready[g], uint32_t(iter) & 1u, err_flag, 12);
          fence_proxy_async_all();
        }
        if (t < n_tiles) {
          const uint8_t* src = ENC ? enc_scratch(g, iter) : in_tiles + size_t(t) * prog.in_tile_stride;
          const uint32_t bytes = uint32_t(prog.in0_nblk) * kBlkBytes;
          mbar_arrive_expect_tx(&in_full[g], bytes * NSPLIT);
          bulk_g2s(act_ptr(g, 0, prog.in0_blk), src + prog.in0_off, bytes, &in_full[g]);
          if (NSPLIT == 2) bulk_g2s(act_ptr(g, 1, prog.in0_blk), src + prog.in0_lo_off, bytes, &in_full[g]);
        } else {
          mbar_arrive(&in_full[g]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&act_ready[g]);   // local; in a CTA pair the peer's helper warp forwards it
    };
#pragma unroll 1
    for (int g = 0; g < NG; ++g)
      if (first_tile(0, g) < n_tiles) tile_prologue(0, g);
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
      for (int l = 0; l < prog.n_layers; ++l) {
        const MlpLayer& L = prog.layers[l];
        int kind = epilogue_kind(L.flags);
        if (NSPLIT == 2 && kind == EK_FINAL_RAW && prog.out_cols == 128) kind = EK_FINAL_RAW_STAGED;
#pragma unroll 1
        for (int g = 0; g < NG; ++g) {
          const long long t0 = first_tile(iter, g);
          if (t0 >= n_tiles) continue;
          const long long t = t0 + cta_rank;         // may be one past the end in the last pair: rows masked, protocol kept
          const bool have_tile = t < n_tiles;
          const long long grow = t * kTileM + row_in_tile;
          const uint32_t act_hi = smem_u32(act_ptr(g, 0, 0));
          const uint32_t act_lo = smem_u32(act_ptr(g, NSPLIT - 1, 0));
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(g * 256);
          float alpha = (g == 0) ? alpha_s0 : alpha_s1;
          float rgb[3] = {0.0f, 0.0f, 0.0f};
          if (l == 0) alpha = 0.0f;
          mbar_wait(&acc_full[g], (acc_phase >> g) & 1u, err_flag, 5);
          acc_phase ^= (1u << g);
          tc_fence_after();
          if (lane == 0 && (e == 0 || e == EW - 1)) tr(1 + 2 * g + (e != 0), g, l, 3);
          // last layer retired -> both bulk copies of this tile's input image were consumed long ago: the encoder
          // warp may overwrite that buffer (monotonic counter: a late reader can not mistake one tile for another)
          if (ENC && l + 1 == prog.n_layers && e == 0 && lane == 0) tiles_done[g] = int(iter) + 1;
          if ((L.flags & LF_LOAD_IN1_AFTER) && e == 0 && lane == 0) {
            if (have_tile) {
              const uint8_t* src = ENC ? enc_scratch(g, iter) : in_tiles + size_t(t) * prog.in_tile_stride;
              mbar_arrive_expect_tx(&in_full[g], kBlkBytes);
              bulk_g2s(act_ptr(g, 0, prog.in1_blk), src + prog.in1_off, kBlkBytes, &in_full[g]);
            } else {
              mbar_arrive(&in_full[g]);
            }
          }
          for (int h = 0; h < L.n_half; ++h) {
            const int c0 = h * 128 + sub * CW;
            switch (kind) {
              case EK_ACT_RELU:
                epilogue_layer<NSPLIT, EK_ACT_RELU>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
              case EK_ACT_RELU_ALPHA:
                epilogue_layer<NSPLIT, EK_ACT_RELU_ALPHA>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
              case EK_ACT_LINEAR:
                epilogue_layer<NSPLIT, EK_ACT_LINEAR>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
              case EK_FINAL_RGB:
                epilogue_layer<NSPLIT, EK_FINAL_RGB>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
              case EK_FINAL_RAW_STAGED:
                epilogue_layer<NSPLIT, EK_FINAL_RAW_STAGED>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
              default:
                epilogue_layer<NSPLIT, EK_FINAL_RAW>(taddr, c0, CW, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
                break;
            }
          }
          if (kind == EK_FINAL_RAW_STAGED) {
            // the tile's [128 x 128] fp32 logits are contiguous in global memory: every warp copies 8 staged rows, one
            // fully coalesced 512-byte row per instruction
            named_bar_sync(1, EW * 32);
#pragma unroll
            for (int i = 0; i < 128 / EW; ++i) {
              const int row = e * (128 / EW) + i;
              const long long gr = t * kTileM + row;
              const uint32_t src = act_hi + ((row < 64) ? 2u : 6u) * kBlkBytes + uint32_t(row & 63) * 512u + ((uint32_t(lane) ^ uint32_t(row & 31)) << 4);
              uint4 q;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(src));
              if (gr < rows) reinterpret_cast<uint4*>(out + gr * 128)[lane] = q;
            }
            named_bar_sync(1, EW * 32);   // staging area is reused by the next tile's hidden activations
          }
          if (L.flags & LF_FINAL_RGB) {
            // The QW warps of a lane quarter hold partial alpha / rgb dot products over their column slices:
            // combine them through shared memory (the slot's hidden blocks are dead once this layer's MMAs
            // have retired) and let the sub == 0 warp write the row.
            if (QW > 1) {
              float4* scratch = reinterpret_cast<float4*>(act_ptr(g, 0, prog.hid_blk0));
              if (sub > 0) scratch[(sub - 1) * kTileM + row_in_tile] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
              // only the QW warps of this lane quarter exchange rows (ids 12..15).  No barrier after the reads: the next
              // writer of these blocks is the next tile's layer-0 epilogue, which cannot start before ALL epilogue warps
              // -- the readers included -- have arrived on act_ready.
              named_bar_sync(12 + quarter, QW * 32);
              if (sub == 0) {
#pragma unroll
                for (int q = 1; q < QW; ++q) {
                  const float4 p = scratch[(q - 1) * kTileM + row_in_tile];
                  rgb[0] += p.x;
                  rgb[1] += p.y;
                  rgb[2] += p.z;
                  alpha += p.w;
                }
              }
            }
            if (sub == 0 && grow < rows) {
              const float ab = prog.side[prog.alpha_b_off];
              const float b0 = prog.side[prog.rgb_b_off], b1 = prog.side[prog.rgb_b_off + 1], b2 = prog.side[prog.rgb_b_off + 2];
              reinterpret_cast<float4*>(out)[grow] = make_float4(rgb[0] + b0, rgb[1] + b1, rgb[2] + b2, alpha + ab);
            }
          }
          if (L.flags & LF_OUT_ACT) fence_proxy_async_smem();
          if (lane == 0 && (e == 0 || e == EW - 1)) tr(1 + 2 * g + (e != 0), g, l, 4);
          if (g == 0) alpha_s0 = alpha; else alpha_s1 = alpha;
          if (l + 1 < prog.n_layers) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[g]);
          } else if (first_tile(iter + 1, g) < n_tiles) {
            tile_prologue(iter + 1, g);   // the slot's next tile starts right away, not after the other slot's last layer
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // nobody leaves while its partner may still signal it
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// Half-pipelined variant for the split-precision sampling net (NSPLIT = 2, one tile per CTA).
//
// One tile does not leave room for a second slot (its hi + lo activations are 128 KB), so the overlap comes from
// inside the layer: every layer's N is produced as two 128-column halves, half 0's epilogue runs while half 1's
// MMAs are still executing, and the next layer's first two K blocks (written by half 0) start while half 1's
// epilogue is still running.  What makes this safe:
//   * accumulators are double buffered in TMEM by a running layer counter (layers G and G+1 never share columns),
//   * hidden activations are still updated in place, so half 0's epilogue may only STORE once the last MMA of the
//     layer that reads hidden blocks 0,1 has retired -- the issuer commits `lo_free` right after those MMAs,
//   * half 1's epilogue writes blocks 2,3 only after all MMAs of the layer have retired (acc_full[1]).
// The next tile's layer 0 needs nothing but its input tile, so a tile's last epilogue overlaps the next tile's
// first MMAs.  Weights are packed N-half outermost for this kernel (adn_set_weights).
template <int CG>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_hp_kernel(const __grid_constant__ MlpProgram prog, const uint8_t* __restrict__ wblob,
              const uint8_t* __restrict__ in_tiles, float* __restrict__ out,
              const long long* __restrict__ rows_dev, long long rows_host, int* err_flag, long long* trace) {
  constexpr int NSPLIT = 2;
  using Ring = RingCfg<NSPLIT, CG>;
  constexpr int NB = MlpCfg<NSPLIT>::kNB;
  constexpr int STAGES = Ring::kStages;
  constexpr int STAGE_BYTES = Ring::kStageBytes;
  constexpr int HALF = kBlkBytes / CG;
  constexpr int EW = 16, QW = 4, CW = 32;
  constexpr int kProducerWarp = 16, kHelperWarp = 17, kMmaWarp = 18;
  constexpr int kBarHalf = 3, kBarStage = 5;   // named barrier ids (1 is used by the epilogue warps)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;                                                   // [NSPLIT][NB] blocks
  uint8_t* ring = act + size_t(NSPLIT) * NB * kBlkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + size_t(STAGES) * STAGE_BYTES);
  uint64_t* w_full = bars;                     // [STAGES]
  uint64_t* w_empty = w_full + STAGES;         // [STAGES]
  uint64_t* peer_full = w_empty + STAGES;      // [STAGES] leader only
  uint64_t* acc_full = peer_full + STAGES;     // [2]  accumulator half complete
  uint64_t* act_ready = acc_full + 2;          // [2]  this CTA's epilogue of that half is done (A blocks written)
  uint64_t* peer_act = act_ready + 2;          // [2]  leader only
  uint64_t* lo_free = peer_act + 2;            // [1]  hidden blocks 0,1 are no longer read by this layer's MMAs
  uint64_t* in_full = lo_free + 1;             // [1]
  uint64_t* peer_in = in_full + 1;             // [1]  leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(peer_in + 1);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const long long n_units = gridDim.x / CG;
  const long long unit = blockIdx.x / CG;
  const long long rows = rows_dev ? *rows_dev : rows_host;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  const int n_layers = prog.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
      mbar_init(&peer_full[s], 1);
    }
    for (int h = 0; h < 2; ++h) {
      mbar_init(&acc_full[h], 1);
      mbar_init(&act_ready[h], EW);
      mbar_init(&peer_act[h], 1);
    }
    mbar_init(lo_free, 1);
    mbar_init(in_full, 1);
    mbar_init(peer_in, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) tmem_alloc_cg<CG>(tmem_slot, 512);
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto act_ptr = [&](int term, int blk) -> uint8_t* { return act + (size_t(term) * NB + blk) * kBlkBytes; };
  auto first_tile = [&](long long iter) -> long long { return (iter * n_units + unit) * CG; };
  (void)trace;

  if (warp == kProducerWarp) {
    // ===================================================================== weight producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter) >= n_tiles) break;
        for (int l = 0; l < n_layers; ++l) {
          const MlpLayer& L = prog.layers[l];
          const int n_st = int(L.n_kb) * int(L.n_half);
          const uint8_t* src = wblob + size_t((blockIdx.x / CG) % prog.w_copies) * prog.w_stride + L.w_off;
          for (int i = 0; i < n_st; ++i) {
            mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 1);
            uint8_t* dst = ring + size_t(stage) * STAGE_BYTES;
            const uint8_t* s0 = src + size_t(i) * (2 * kBlkBytes);
            mbar_arrive_expect_tx(&w_full[stage], 2 * HALF);
            bulk_g2s(dst, s0 + cta_rank * HALF, HALF, &w_full[stage]);
            bulk_g2s(dst + HALF, s0 + kBlkBytes + cta_rank * HALF, HALF, &w_full[stage]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kHelperWarp) {
    // =================================================================== barrier helper (see mlp_umma_kernel)
    int stage = 0;
    uint32_t phase = 0, in_phase = 0, ar_phase[2] = {0, 0};
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter) >= n_tiles) break;
      for (int l = 0; l < n_layers; ++l) {
        const MlpLayer& L = prog.layers[l];
        if (l == 0) {
          mbar_wait(in_full, in_phase, err_flag, 2);
          if (CG == 2) {
            if (leader) mbar_wait(peer_in, in_phase, err_flag, 7);
            else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(peer_in), 0));
          }
          in_phase ^= 1;
          if (leader) named_bar_arrive(kBarHalf, 64);
        }
        bool seen[2] = {l == 0, l == 0};
        for (int nh = 0; nh < L.n_half; ++nh) {
          for (int kb = 0; kb < L.n_kb; ++kb) {
            const int h = int(L.a_blk[kb]) >> 1;   // hidden blocks 0,1 <- half 0 of the previous layer, 2,3 <- half 1
            if (!seen[h]) {
              seen[h] = true;
              mbar_wait(&act_ready[h], ar_phase[h], err_flag, 3);
              if (CG == 2) {
                if (leader) mbar_wait(&peer_act[h], ar_phase[h], err_flag, 9);
                else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_act[h]), 0));
              }
              ar_phase[h] ^= 1;
              if (leader) named_bar_arrive(kBarHalf + h, 64);
            }
            mbar_wait(&w_full[stage], phase, err_flag, 4);
            if (CG == 2) {
              if (leader) mbar_wait(&peer_full[stage], phase, err_flag, 8);
              else if (lane == 0) mbar_arrive_remote(mapa_shared(smem_u32(&peer_full[stage]), 0));
            }
            if (leader) named_bar_arrive(kBarStage + stage, 64);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ========================================================================== MMA issuer (leader CTA only)
    if (leader) {
      constexpr uint32_t idesc128 = make_idesc_bf16(128 * CG, 128);
      int stage = 0;
      uint32_t gl = 0;   // running layer counter: selects the TMEM accumulator buffer
      const uint64_t desc_hi = make_desc_sw128(0) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo_const = uint32_t(make_desc_sw128(0) & 0xFFFF0000ull);
      auto lo_of = [&](uint32_t addr) -> uint32_t { return desc_lo_const | (addr >> 4); };
      const uint32_t ring_lo = lo_of(smem_u32(ring));
      const uint32_t act_lo0 = lo_of(smem_u32(act));
      auto desc = [&](uint32_t lo) -> uint64_t { return desc_hi | uint64_t(lo); };
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter) >= n_tiles) break;
        for (int l = 0; l < n_layers; ++l, ++gl) {
          const MlpLayer& L = prog.layers[l];
          const bool final_layer = (l + 1 == n_layers);
          uint32_t blks = 0;
#pragma unroll
          for (int kb = 0; kb < 6; ++kb) blks |= uint32_t(L.a_blk[kb] & 15) << (4 * kb);
          if (l == 0) {
            named_bar_sync(kBarHalf, 64);   // input tile landed (both CTAs)
            tc_fence_after();
          }
          bool seen[2] = {l == 0, l == 0};
          const uint32_t d_layer = tmem_base + (gl & 1u) * 256u;
          for (int nh = 0; nh < L.n_half; ++nh) {
            for (int kb = 0; kb < L.n_kb; ++kb) {
              const uint32_t blk = (blks >> (4 * kb)) & 15u;
              const int h = int(blk >> 1);
              if (!seen[h]) {
                seen[h] = true;
                named_bar_sync(kBarHalf + h, 64);   // previous layer's half h has been published by both CTAs
                tc_fence_after();
              }
              const uint32_t a_hi = act_lo0 + blk * (kBlkBytes >> 4);
              const uint32_t a_lo = a_hi + uint32_t(NB) * (kBlkBytes >> 4);
              named_bar_sync(kBarStage + stage, 64);
              tc_fence_after();
              const uint32_t b_hi = ring_lo + uint32_t(stage) * (STAGE_BYTES >> 4);
              const uint32_t b_lo = b_hi + (HALF >> 4);
              const uint32_t d = d_layer + uint32_t(nh * 128);
              if (elect_one()) {
                const int nk = L.k_cnt[kb];   // zero-padded tail columns of an input block are not multiplied
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (k < nk) {
                    umma_bf16_cg<CG>(d, desc(a_hi + 2 * k), desc(b_hi + 2 * k), idesc128, (kb > 0 || k > 0) ? 1u : 0u);
                    umma_bf16_cg<CG>(d, desc(a_lo + 2 * k), desc(b_hi + 2 * k), idesc128, 1u);
                    umma_bf16_cg<CG>(d, desc(a_hi + 2 * k), desc(b_lo + 2 * k), idesc128, 1u);
                  }
                }
                umma_commit_cg<CG>(&w_empty[stage]);
                // the last MMAs of this layer that read hidden blocks 0,1 have been issued: half 0's epilogue may store
                if (!final_layer && nh == L.n_half - 1 && kb == 1) umma_commit_cg<CG>(lo_free);
              }
              __syncwarp();
              if (++stage == STAGES) stage = 0;
            }
            if (elect_one()) umma_commit_cg<CG>(&acc_full[nh]);
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ============================================================================ epilogue (16 warps, one tile)
    const int e = warp;                            // 0..15
    const int quarter = warp & 3;
    const int sub = e >> 2;                        // 32-column slice of each N half
    const int row_in_tile = quarter * 32 + lane;
    uint32_t acc_ph[2] = {0, 0}, lo_ph = 0, gl = 0;
    const uint32_t act_hi = smem_u32(act_ptr(0, 0));
    const uint32_t act_lo = smem_u32(act_ptr(1, 0));
    float alpha = 0.f, rgb[3] = {0.f, 0.f, 0.f};
    for (long long iter = 0;; ++iter) {
      const long long t0 = first_tile(iter);
      if (t0 >= n_tiles) break;
      const long long t = t0 + cta_rank;
      const bool have_tile = t < n_tiles;
      const long long grow = t * kTileM + row_in_tile;
      if (e == 0 && lane == 0) {
        // safe: this warp has seen acc_full of the previous tile's last half, i.e. every MMA reading blocks 0,1 retired
        if (have_tile) {
          const uint8_t* src = in_tiles + size_t(t) * prog.in_tile_stride;
          const uint32_t bytes = uint32_t(prog.in0_nblk) * kBlkBytes;
          mbar_arrive_expect_tx(in_full, bytes * NSPLIT);
          bulk_g2s(act_ptr(0, prog.in0_blk), src + prog.in0_off, bytes, in_full);
          bulk_g2s(act_ptr(1, prog.in0_blk), src + prog.in0_lo_off, bytes, in_full);
        } else {
          mbar_arrive(in_full);
        }
      }
      for (int l = 0; l < n_layers; ++l, ++gl) {
        const MlpLayer& L = prog.layers[l];
        const bool final_layer = (l + 1 == n_layers);
        int kind = epilogue_kind(L.flags);
        if (kind == EK_FINAL_RAW && prog.out_cols == 128) kind = EK_FINAL_RAW_STAGED;
        const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + (gl & 1u) * 256u;
        for (int h = 0; h < L.n_half; ++h) {
          mbar_wait(&acc_full[h], acc_ph[h], err_flag, 5);
          acc_ph[h] ^= 1;
          tc_fence_after();
          const int c = h * 128 + sub * CW;
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tc_wait_ld();
          if (h == 0 && !final_layer) {   // in-place hazard: hidden blocks 0,1 may still feed this layer's half-1 MMAs
            mbar_wait(lo_free, lo_ph, err_flag, 6);
            lo_ph ^= 1;
          }
          switch (kind) {
            case EK_ACT_RELU:
              epilogue_chunk<NSPLIT, EK_ACT_RELU>(r, c, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
              break;
            case EK_ACT_LINEAR:
              epilogue_chunk<NSPLIT, EK_ACT_LINEAR>(r, c, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
              break;
            case EK_FINAL_RAW_STAGED:
              epilogue_chunk<NSPLIT, EK_FINAL_RAW_STAGED>(r, c, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
              break;
            default:
              epilogue_chunk<NSPLIT, EK_FINAL_RAW>(r, c, L, prog, act_hi, act_lo, row_in_tile, grow, rows, out, alpha, rgb);
              break;
          }
          if (!final_layer) {
            fence_proxy_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&act_ready[h]);
          }
        }
        if (final_layer) {
          // every epilogue warp meets here once per tile: the staged logits are complete (and, for tiny test nets,
          // nobody is still reading an accumulator buffer that the next tile's first layers will overwrite)
          tc_fence_before();
          named_bar_sync(1, EW * 32);
          if (kind == EK_FINAL_RAW_STAGED) {
#pragma unroll
            for (int i = 0; i < 128 / EW; ++i) {
              const int row = e * (128 / EW) + i;
              const long long gr = t * kTileM + row;
              const uint32_t src = act_hi + ((row < 64) ? 2u : 6u) * kBlkBytes + uint32_t(row & 63) * 512u + ((uint32_t(lane) ^ uint32_t(row & 31)) << 4);
              uint4 q;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(src));
              if (gr < rows) reinterpret_cast<uint4*>(out + gr * 128)[lane] = q;
            }
            named_bar_sync(1, EW * 32);   // staging area = hidden blocks 2,3, rewritten by the next tile's layer-0 epilogue
          }
        }
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// fp32 feature rows [rows, n_feat] -> packed bf16 (hi / lo) SWIZZLE_128B tile blocks.
// One thread per (row, block, 16-byte chunk): 8 consecutive source columns.
__global__ void pack_rows_kernel(const float* __restrict__ x, long long rows_host, const long long* __restrict__ rows_dev,
                                 int n_feat, const __grid_constant__ InputLayout lay, uint8_t* __restrict__ tiles) {
  const long long rows = rows_dev ? *rows_dev : rows_host;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  const long long total = n_tiles * kTileM * lay.n_blk * 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int chunk = int(i & 7);
    const long long rb = i >> 3;
    const int blk = int(rb % lay.n_blk);
    const long long row = rb / lay.n_blk;
    const int r = int(row & (kTileM - 1));
    const long long t = row >> 7;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = chunk * 8 + j;
      v[j] = (row < rows && kk < lay.valid[blk]) ? x[row * n_feat + lay.src_col0[blk] + kk] : 0.0f;
    }
    uint4 hi;
    hi.x = pack_bf16x2(v[0], v[1]);
    hi.y = pack_bf16x2(v[2], v[3]);
    hi.z = pack_bf16x2(v[4], v[5]);
    hi.w = pack_bf16x2(v[6], v[7]);
    const uint32_t off = sw128_offset(uint32_t(r), uint32_t(chunk * 8));
    uint8_t* tb = tiles + size_t(t) * lay.tile_stride;
    *reinterpret_cast<uint4*>(tb + lay.dst_off_hi[blk] + off) = hi;
    if (lay.nsplit == 2) {
      const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
      float l[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        l[2 * e + 0] = v[2 * e + 0] - __uint_as_float(hw[e] << 16);
        l[2 * e + 1] = v[2 * e + 1] - __uint_as_float(hw[e] & 0xFFFF0000u);
      }
      uint4 lo;
      lo.x = pack_bf16x2(l[0], l[1]);
      lo.y = pack_bf16x2(l[2], l[3]);
      lo.z = pack_bf16x2(l[4], l[5]);
      lo.w = pack_bf16x2(l[6], l[7]);
      *reinterpret_cast<uint4*>(tb + lay.dst_off_lo[blk] + off) = lo;
    }
  }
}

// -------------------------------------------------------------------------------------------------
template <int NSPLIT, int NG, int CG, bool ENC = false>
static cudaError_t launch_mlp_t(const MlpProgram& prog, const uint8_t* wblob, const uint8_t* in_tiles,
                                float* out, const long long* rows_dev, long long rows_host, int* err_flag, int num_sms,
                                cudaStream_t stream, long long* trace, const EncodeParams* encp = nullptr) {
  static unsigned long long attr_done = 0;   // per device (the attribute is per device)
  const size_t smem = mlp_smem_layout_bytes<NSPLIT, NG, CG>();
  auto kernel = mlp_umma_kernel<NSPLIT, NG, CG, ENC>;
  const EncodeParams enc = encp ? *encp : EncodeParams{};
  {
    cudaError_t e = set_max_dyn_smem_once(reinterpret_cast<const void*>(kernel), int(smem), &attr_done);
    if (e != cudaSuccess) return e;
  }
  int grid = (num_sms / CG) * CG;   // persistent: one CTA (or CTA pair) per SM (pair)
  if (!rows_dev) {
    const long long n_tiles = (rows_host + kTileM - 1) / kTileM;
    const long long need = ((n_tiles + NG * CG - 1) / (NG * CG)) * CG;
    if (need < grid) grid = int(need < CG ? CG : need);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(ENC ? kMlpEncThreads : kMlpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, trace, enc);
}

template <int CG>
static cudaError_t launch_hp_t(const MlpProgram& prog, const uint8_t* wblob, const uint8_t* in_tiles, float* out,
                               const long long* rows_dev, long long rows_host, int* err_flag, int num_sms, cudaStream_t stream,
                               long long* trace) {
  static unsigned long long attr_done = 0;
  const size_t smem = mlp_smem_layout_bytes<2, 1, CG>();
  auto kernel = mlp_hp_kernel<CG>;
  {
    cudaError_t e = set_max_dyn_smem_once(reinterpret_cast<const void*>(kernel), int(smem), &attr_done);
    if (e != cudaSuccess) return e;
  }
  int grid = (num_sms / CG) * CG;
  if (!rows_dev) {
    const long long n_tiles = (rows_host + kTileM - 1) / kTileM;
    const long long need = ((n_tiles + CG - 1) / CG) * CG;
    if (need < grid) grid = int(need < CG ? CG : need);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(kMlpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, trace);
}

size_t mlp_enc_scratch_bytes(int num_sms) { return size_t(num_sms) * 2 * 2 * (2 * kBlkBytes); }

cudaError_t launch_mlp(int nsplit, int ng, int cg, const MlpProgram& prog, const uint8_t* wblob,
                       const uint8_t* in_tiles, float* out, const long long* rows_dev, long long rows_host, int* err_flag,
                       int num_sms, cudaStream_t stream, long long* trace, const EncodeParams* enc) {
  if (enc) {   // shading net with the fused input encoder (CTA pairs only)
    if (nsplit != 1 || cg != 2) return cudaErrorInvalidValue;
    return launch_mlp_t<1, 2, 2, true>(prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, num_sms, stream, trace, enc);
  }
  // split precision (sampling net): half-pipelined single-tile kernel; weights packed N-half outermost
  if (nsplit == 2) {
    if (cg == 2) return launch_hp_t<2>(prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, num_sms, stream, trace);
    return launch_hp_t<1>(prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, num_sms, stream, trace);
  }
  if (cg == 2) return launch_mlp_t<1, 2, 2>(prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, num_sms, stream, trace);
  return launch_mlp_t<1, 2, 1>(prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, num_sms, stream, trace);
}

cudaError_t launch_pack_rows(const float* x, long long rows, const long long* rows_dev, int n_feat, const InputLayout& lay,
                             uint8_t* tiles, cudaStream_t stream) {
  long long work = ((rows + kTileM - 1) / kTileM) * kTileM * lay.n_blk * 8;
  int grid = int((work + 255) / 256);
  if (grid < 1) grid = 1;
  if (grid > 148 * 16) grid = 148 * 16;
  pack_rows_kernel<<<grid, 256, 0, stream>>>(x, rows, rows_dev, n_feat, lay, tiles);
  return cudaGetLastError();
}

}  // namespace adn
