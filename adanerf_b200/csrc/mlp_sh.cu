// Shading MLP (NeRF 8x256, skip 4, view branch; src/models.py:254-277 of the reference) on tcgen05 / TMEM: CTA pairs,
// two tile slots per CTA that alternate layer by layer (slot 0's epilogue runs under slot 1's MMAs), N = 256 MMAs.
//
// How it differs from mlp_umma_kernel<1,2,2> (round 1), and what was measured on the way (profiles/trace_sh.py):
//   * Tile inputs (position block at layers 0 and 5, view block at the last layer) are A operands that travel through the
//     same FIFO ring as the weights, fetched by the producer right before the step that needs them.  No resident input
//     block per slot: the 32 KB go to the ring (5 stages of 16 KB instead of 4; the round-1 issue loop ran at 82 % of the
//     pipe rate waiting for weights), and the next tile's input is prefetched instead of costing ~3000 cycles per tile.
//   * The next layer of a slot starts before its epilogue has finished: the epilogue first drains the whole accumulator
//     into registers (two TMEM loads in flight) and says so (`drained`), then publishes the activations in two halves
//     (`act0`: columns 0-127 = K blocks 0, 1 of the next layer; `act1`: the rest).  The issuer needs `drained` + `act0`
//     for the next layer's first K block and `act1` only two K blocks later: the MMA -> epilogue -> MMA chain that left
//     the tensor pipe idle ~20 % of a layer pair in round 1 is off the critical path.
//   * Both CTAs of a pair arrive on the LEADER's barriers (the peer's epilogue warps with relaxed remote arrives after their
//     proxy fence, its helper warp forwards "my share of the stage has landed"): one barrier per dependency instead of a
//     local + forwarded pair.  A helper warp waits for everything a step needs (weight stage, input stage, the slot's
//     epilogue barriers) at once, one mbarrier per lane, ahead of the issuer, and hands over through a named barrier: an
//     mbarrier wait in the issuing warp itself costs 500-700 cycles per 512-cycle step (measured with two issuer warps).
//   * Biases and head vectors are read from a shared-memory copy of the side parameters with warp-uniform 16-byte loads
//     (the register-indexed constant loads of the round-1 epilogue cost 0.5 ms of 5.0), fp32 exact.
//   * Tried and dropped (kept in git history): N-half pipelining with one weight stage feeding both slots.  It halves the
//     L2 -> SM weight traffic, but its N = 128 MMAs read A (4 KB) and B (2 KB) from shared memory every 64 tensor cycles
//     -- 96 B / clk against the ~64 B / clk the tensor core gets (exactly what N = 256 needs): with an EMPTY epilogue
//     the 580 MMAs of a tile pair took 52.7 k cycles (91 per MMA), 65 k with the real one -- no better than round 1.
//     A fully specialised epilogue (immediate bias operands, 160 KB of code) missed the instruction cache on every event.
//
// Roles (19 warps): 16 epilogue warps (all of them serve both slots), producer, barrier helper, MMA issuer (leader CTA).
// Issue order:  for tile group:  for layer:  for slot g:  for step (one ring stage of weights).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "mlp_umma.cuh"
#include "ptx.cuh"

namespace adn {

namespace {

#ifndef ADN_SH_DIAG
#define ADN_SH_DIAG 0   // timing experiments only (wrong results): 1 = epilogue events do nothing but signal
#endif
constexpr int kShThreads = 608;            // 19 warps
constexpr int kShStages = 5;               // ring stages per CTA
constexpr int kShStageBytes = kBlkBytes;   // weights: one K block x this CTA's 128 of 256 N rows (N = 128 layers: two K blocks x 64
                                           // rows); or one tile-input block
constexpr int kShNB = 4;                   // hidden activation blocks per slot

enum : int { SK_RELU = 0, SK_RELU_ALPHA = 1, SK_LINEAR = 2, SK_RGB = 3 };

__device__ __forceinline__ uint32_t sh_pack(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t sh_pack_relu(float a, float b) {
  uint32_t d;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
  return d;
}
// Side-table read (constant after the kernel prologue): a plain asm so the compiler may schedule it freely.
__device__ __forceinline__ float4 sh_side4(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// 32 accumulator columns [c, c + 32) of one row: bias (+ ReLU, + head partial sums), packed to bf16 pairs.
// Generic over the layer: biases and head vectors come from the shared-memory copy of MlpProgram::side through
// warp-uniform (broadcast) 16-byte loads.  (A version specialised per layer with immediate constant operands was
// 160 KB of code, missed the instruction cache on every event and ran at a third of the speed.)
template <int KIND>
__device__ __forceinline__ void sh_chunk(const uint32_t (&r)[32], uint32_t side_s, int boff, int c, uint32_t (&p)[16], float& alpha,
                                         float (&rgb)[3]) {
  float v[32];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 b = sh_side4(side_s + uint32_t(boff + 4 * j) * 4u);
    v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) + b.x;
    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z;
    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
  }
  if (KIND == SK_RELU_ALPHA || KIND == SK_RGB) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (KIND == SK_RELU_ALPHA) {   // alpha_linear on the fp32 post-activation row (models.py:264), four partial sums
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 w = sh_side4(side_s + uint32_t(kShAlphaW + c + 4 * j) * 4u);
      a0 = fmaf(v[4 * j + 0], w.x, a0);
      a1 = fmaf(v[4 * j + 1], w.y, a1);
      a2 = fmaf(v[4 * j + 2], w.z, a2);
      a3 = fmaf(v[4 * j + 3], w.w, a3);
    }
    alpha += (a0 + a1) + (a2 + a3);
  }
  if (KIND == SK_RGB) {          // rgb_linear (models.py:272) on the fp32 row
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 w = sh_side4(side_s + uint32_t(kShRgbW + k * 128 + c + 4 * j) * 4u);
        a0 = fmaf(v[4 * j + 0], w.x, a0);
        a1 = fmaf(v[4 * j + 1], w.y, a1);
        a2 = fmaf(v[4 * j + 2], w.z, a2);
        a3 = fmaf(v[4 * j + 3], w.w, a3);
      }
      rgb[k] += (a0 + a1) + (a2 + a3);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) p[j] = (KIND == SK_RELU) ? sh_pack_relu(v[2 * j], v[2 * j + 1]) : sh_pack(v[2 * j], v[2 * j + 1]);
  }
}

template <int KIND>
__device__ __forceinline__ void sh_store(const uint32_t (&p)[16], uint32_t st_row, uint32_t unit0, uint32_t rx) {
#pragma unroll
  for (int q = 0; q < 4; ++q) st_shared_v4(st_row + (((unit0 + q) ^ rx) << 4), p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
}

}  // namespace

__global__ void __launch_bounds__(kShThreads, 1)
mlp_sh_kernel(const __grid_constant__ MlpProgram prog, const uint8_t* __restrict__ wblob, const uint8_t* __restrict__ in_tiles,
              float* __restrict__ out, const long long* __restrict__ rows_dev, long long rows_host, int* err_flag, long long* trace) {
  constexpr int STAGES = kShStages, STAGE_BYTES = kShStageBytes, NB = kShNB;
  constexpr int EW = 16, QW = 4;
  // named barriers helper -> issuer, one per step, cyclic: the helper can run at most STAGES steps ahead of the issuer (a
  // step's weight stage is loaded only after the stage STAGES ring positions earlier was released), so STAGES + 1 ids
  // are never reused before the issuer has consumed them
  constexpr int kStepBars = kShStages + 1;
  constexpr int kProducerWarp = 16, kHelperWarp = 17, kMmaWarp = 18;   // highest warp id: favoured by the issue arbiter
  // named barriers 1..6: helper -> issuer, one per step (cyclic); 12..15: rgb exchange of the epilogue warps of a lane quarter

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;                                        // [2 slots][NB] hidden blocks
  uint8_t* ring = act + size_t(2) * NB * kBlkBytes;           // [STAGES]
  float* side_s = reinterpret_cast<float*>(ring + size_t(STAGES) * STAGE_BYTES);   // copy of prog.side (biases, head vectors)
  uint64_t* bars = reinterpret_cast<uint64_t*>(side_s + kSideFloats);
  uint64_t* w_full = bars;                   // [STAGES] this CTA's share of the stage has landed (leader: and the peer's)
  uint64_t* w_empty = w_full + STAGES;       // [STAGES] the MMAs reading the stage have retired (both CTAs)
  uint64_t* acc_full = w_empty + STAGES;     // [g] all MMAs of the slot's layer have retired
  uint64_t* drained = acc_full + 2;          // [g] leader only: both CTAs' epilogue warps hold the accumulator in registers
  uint64_t* act_half = drained + 2;          // [2 g + hh] leader only: columns [128 hh, +128) of the layer are written (both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_half + 4);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = (cta_rank == 0);
  const long long n_units = gridDim.x / 2;
  const long long unit = blockIdx.x / 2;
  const long long rows = rows_dev ? *rows_dev : rows_host;
  const long long n_tiles = (rows + kTileM - 1) / kTileM;
  const int n_layers = prog.n_layers;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&w_full[s], leader ? 2 : 1);   // producer's expect_tx arrive (+ in the leader: the peer's forwarded "landed")
      mbar_init(&w_empty[s], 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&acc_full[g], 1);
      mbar_init(&drained[g], 2 * EW);          // leader only: the epilogue warps of BOTH CTAs arrive here
      mbar_init(&act_half[2 * g], 2 * EW);
      mbar_init(&act_half[2 * g + 1], 2 * EW);
    }
    mbar_fence_init();
  }
  for (int i = threadIdx.x; i < kSideFloats; i += kShThreads) side_s[i] = prog.side[i];
  if (warp == kMmaWarp) tmem_alloc_cg<2>(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // hidden activation block `blk` (program numbering: 1..4; 0 is the tile input, which lives in the ring) of slot g
  auto act_ptr = [&](int g, int blk) -> uint8_t* { return act + (size_t(g) * NB + (blk - 1)) * kBlkBytes; };
  // first tile of the pair's tile group (iter, slot); CTA `cta_rank` owns tile first + cta_rank
  auto first_tile = [&](long long iter, int g) -> long long { return ((iter * n_units + unit) * 2 + g) * 2; };
  // Optional timeline (debug, option "trace"): CTA 0 records (clock, code) pairs; region r = words [8192 r, 8192 (r+1)),
  // word 0 = count; code = slot << 16 | layer << 8 | event.
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
#ifdef ADN_SH_CKPT
  // debug: progress markers of CTA 0 / CTA 1 in host-mapped memory behind the watchdog flag (readable after a fault)
  auto ckpt = [&](int slot, int code) {
    if (blockIdx.x < 2) reinterpret_cast<volatile int*>(err_flag)[1 + slot + 5 * int(blockIdx.x)] = code;
  };
#else
  auto ckpt = [&](int, int) {};
#endif
  if (tracing && threadIdx.x == 0) {   // region 7: kernel start / end in SM clocks and in ns (the real SM frequency)
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[7 * 8192 + 0] = clock64();
    trace[7 * 8192 + 1] = gt;
  }

  if (warp == kProducerWarp) {
    // ===================================================================== producer (weights and tile inputs)
    // Weight stage of a step: N = 256 layers: K block s, this CTA's N half ([128 x 64] tile `rank` of the K block's pair);
    // N = 128 layers: K blocks 2s, 2s+1, this CTA's rows [64 rank, +64) of each [128 x 64] tile.  A step whose K block is
    // the tile input is followed by a stage holding the slot's input block.
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter, 0) >= n_tiles) break;
        const bool active1 = first_tile(iter, 1) < n_tiles;
        for (int l = 0; l < n_layers; ++l) {
          const MlpLayer& L = prog.layers[l];
          const int n_steps = prog.sh_steps[l];
          for (int g = 0; g < 2; ++g) {
            if (g == 1 && !active1) continue;
            for (int i = 0; i < n_steps; ++i) {
              const uint32_t w = prog.sh_sched[l][i];
              tr(5, stage, l, 8);
              ckpt(0, 1000 * int(iter) + 100 * l + 10 * g + i);
              mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 1);
              tr(5, stage, l, 9);
              uint8_t* dst = ring + size_t(stage) * STAGE_BYTES;
              if (w & (1u << 14)) {   // N = 128
                const uint8_t* src = wblob + L.w_off + size_t(2 * i) * kBlkBytes + cta_rank * (kBlkBytes / 2);
                const int nkb = (w & (1u << 8)) ? 2 : 1;
                mbar_arrive_expect_tx(&w_full[stage], uint32_t(nkb) * (kBlkBytes / 2));
                bulk_g2s(dst, src, kBlkBytes / 2, &w_full[stage]);
                if (nkb == 2) bulk_g2s(dst + kBlkBytes / 2, src + kBlkBytes, kBlkBytes / 2, &w_full[stage]);
              } else {   // N = 256: K blocks 2i (and 2i+1 in the next stage), this CTA's N half of each
                const uint8_t* src = wblob + L.w_off + size_t(4 * i + cta_rank) * kBlkBytes;
                mbar_arrive_expect_tx(&w_full[stage], kBlkBytes);
                bulk_g2s(dst, src, kBlkBytes, &w_full[stage]);
                if (w & (1u << 8)) {
                  advance();
                  mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 1);
                  mbar_arrive_expect_tx(&w_full[stage], kBlkBytes);
                  bulk_g2s(ring + size_t(stage) * STAGE_BYTES, src + 2 * kBlkBytes, kBlkBytes, &w_full[stage]);
                }
              }
              advance();
              if (w & (1u << 16)) {   // this step's K block is the slot's tile input (such a step holds one K block)
                mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 2);
                const long long t = first_tile(iter, g) + cta_rank;
                if (t < n_tiles) {
                  const uint32_t off = (w & (1u << 19)) ? prog.in1_off : prog.in0_off;
                  mbar_arrive_expect_tx(&w_full[stage], kBlkBytes);
                  bulk_g2s(ring + size_t(stage) * STAGE_BYTES, in_tiles + size_t(t) * prog.in_tile_stride + off, kBlkBytes, &w_full[stage]);
                } else {
                  mbar_arrive(&w_full[stage]);   // no such tile (last, odd pair): rows are masked, the protocol is kept
                }
                advance();
              }
            }
          }
        }
      }
    }
  } else if (warp == kHelperWarp) {
    // =================================================================== barrier helper (both CTAs)
    // Walks the issue schedule ahead of the issuer and does its mbarrier waiting for it (an mbarrier round trip costs a
    // few hundred cycles even when the phase completed long ago; 4 MMAs of a step are 512 tensor cycles).  What a step
    // waits for, one barrier per lane, all at once: lane 0: the slot's accumulator drained; lanes 1, 2: columns 0-127 /
    // 128-255 of the previous layer written (each the first time the layer needs it); lane 3: the tile-input stage;
    // lane 4: the weight stage.  Leader: the result is a named-barrier arrival (ids 1..6, cyclic per step).  Peer: lanes 3, 4 forward "my share has
    // landed" to the leader's w_full with relaxed remote arrives (they have acquired the local barrier first; a release
    // at cluster scope would cost a MEMBAR.ALL.GPU).
    int stage = 0, nb = 0;
    uint32_t phase = 0;
    uint32_t dep_phase = 0;   // bit 3 g + k: k = 0: drained[g]; k = 1, 2: act_half[2 g + 0 / 1]
    auto advance = [&]() {
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
      const bool active1 = first_tile(iter, 1) < n_tiles;
      for (int l = 0; l < n_layers; ++l) {
        const int n_steps = prog.sh_steps[l];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (g == 1 && !active1) continue;
          for (int i = 0; i < n_steps; ++i) {
            const uint32_t w = prog.sh_sched[l][i];
            const uint32_t need = leader ? ((w >> 9) & 7u) : 0u;
            const int w_stage = stage;
            const uint32_t w_par = phase;
            advance();
            // second stage of the step: the tile input, or (N = 256 layers) the step's second K block
            const bool input = (w & (1u << 16)) != 0;
            const bool second = input || ((w & (1u << 8)) && !(w & (1u << 14)));
            const int in_stage = stage;
            const uint32_t in_par = phase;
            if (second) advance();
            const bool mine = lane < 3 ? ((need >> lane) & 1u) != 0 : (lane == 3 ? second : lane == 4);
            uint64_t* bar = lane == 0 ? &drained[g] : (lane < 3 ? &act_half[2 * g + lane - 1] : (lane == 3 ? &w_full[in_stage] : &w_full[w_stage]));
            const uint32_t parity = lane < 3 ? ((dep_phase >> (3 * g + lane)) & 1u) : (lane == 3 ? in_par : w_par);
            if (lane == 0) tr(6, g, l, 10);
            mbar_wait_lanes(bar, parity, mine, err_flag, 3);
            if (lane == 0) tr(6, g, l, 11);
            if (!leader && ((lane == 3 && second) || lane == 4)) mbar_arrive_remote(mapa_shared(smem_u32(bar), 0));
            __syncwarp();
            dep_phase ^= need << (3 * g);
            if (leader) named_bar_arrive(1 + nb, 64);
            if (++nb == kStepBars) nb = 0;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ========================================================================== MMA issuer (leader CTA only)
    // The whole warp walks the schedule converged (descriptors in uniform registers); one elected lane issues.
    if (leader) {
      constexpr uint32_t idesc256 = make_idesc_bf16(256, 256);
      constexpr uint32_t idesc128 = make_idesc_bf16(256, 128);
      const uint64_t desc_hi = make_desc_sw128(0) & 0xFFFFFFFF00000000ull;
      const uint32_t desc_lo_const = uint32_t(make_desc_sw128(0) & 0xFFFF0000ull);
      auto lo_of = [&](uint32_t addr) -> uint32_t { return desc_lo_const | (addr >> 4); };
      const uint32_t ring_lo = lo_of(smem_u32(ring));
      const uint32_t act_lo0 = lo_of(smem_u32(act));
      auto desc = [&](uint32_t lo) -> uint64_t { return desc_hi | uint64_t(lo); };
      int stage = 0, nb = 0;
      auto advance = [&]() {
        if (++stage == STAGES) stage = 0;
      };
      for (long long iter = 0;; ++iter) {
        if (first_tile(iter, 0) >= n_tiles) break;
        const bool active1 = first_tile(iter, 1) < n_tiles;
        for (int l = 0; l < n_layers; ++l) {
          const int n_steps = prog.sh_steps[l];
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (g == 1 && !active1) continue;
            const uint32_t d = tmem_base + uint32_t(g * 256);
            const uint32_t act_lo = act_lo0 + uint32_t(g * NB) * (kBlkBytes >> 4);
            for (int i = 0; i < n_steps; ++i) {
              const uint32_t w = prog.sh_sched[l][i];
              const int w_stage = stage;
              advance();
              const bool input = (w & (1u << 16)) != 0;
              const bool second = input || ((w & (1u << 8)) && !(w & (1u << 14)));
              const int in_stage = stage;   // the step's second stage: tile input, or (N = 256) its second K block
              if (second) advance();
              if (lane == 0) tr(0, i, l, 0);
              named_bar_sync(1 + nb, 64);   // helper: weights (and input) landed in both CTAs, the slot's dependencies are met
              if (++nb == kStepBars) nb = 0;
              tc_fence_after();
              if (lane == 0) tr(0, g, l, (w & (1u << 15)) ? 1 : 6);
              const uint32_t acc = ((w >> 15) & 1u) ^ 1u;   // first step of the layer: overwrite
              const uint32_t b = ring_lo + uint32_t(w_stage) * (STAGE_BYTES >> 4);
              const uint32_t a0 = input ? ring_lo + uint32_t(in_stage) * (STAGE_BYTES >> 4) : act_lo + ((w & 15u) - 1u) * (kBlkBytes >> 4);
              if (elect_one()) {
                if (w & (1u << 14)) {   // N = 128: two K blocks per stage, this CTA holds 64 B rows of each
                  const uint32_t a1 = act_lo + (((w >> 4) & 15u) - 1u) * (kBlkBytes >> 4);
                  umma_bf16_cg<2>(d, desc(a0), desc(b), idesc128, acc);
                  umma_bf16_cg<2>(d, desc(a0 + 2), desc(b + 2), idesc128, 1u);
                  if (!(w & (1u << 19))) {   // the view block carries 27 features: two K steps
                    umma_bf16_cg<2>(d, desc(a0 + 4), desc(b + 4), idesc128, 1u);
                    umma_bf16_cg<2>(d, desc(a0 + 6), desc(b + 6), idesc128, 1u);
                  }
                  if (w & (1u << 8)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_cg<2>(d, desc(a1 + 2 * k), desc(b + (kBlkBytes >> 5) + 2 * k), idesc128, 1u);
                  }
                } else {
                  umma_bf16_cg<2>(d, desc(a0), desc(b), idesc256, acc);
#pragma unroll
                  for (int k = 1; k < 4; ++k) umma_bf16_cg<2>(d, desc(a0 + 2 * k), desc(b + 2 * k), idesc256, 1u);
                  if (w & (1u << 8)) {   // second K block: next ring stage
                    umma_commit_cg<2>(&w_empty[w_stage]);   // the first stage is free as soon as its four MMAs have retired
                    const uint32_t a1 = act_lo + (((w >> 4) & 15u) - 1u) * (kBlkBytes >> 4);
                    const uint32_t b1 = ring_lo + uint32_t(in_stage) * (STAGE_BYTES >> 4);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16_cg<2>(d, desc(a1 + 2 * k), desc(b1 + 2 * k), idesc256, 1u);
                  }
                }
                if (!((w & (1u << 8)) && !(w & (1u << 14)))) umma_commit_cg<2>(&w_empty[w_stage]);
                if (second) umma_commit_cg<2>(&w_empty[in_stage]);
                if (w & (1u << 12)) umma_commit_cg<2>(&acc_full[g]);
              }
              __syncwarp();
              if (lane == 0 && (w & (1u << 12))) tr(0, g, l, 2);
            }
          }
        }
      }
    }
  } else {
    // ============================================================================ epilogue (16 warps, both slots)
    // Events in the issuers' completion order: for layer, for slot.  Per event every warp takes two 32-column chunks of the
    // [128 x 256] accumulator -- columns [32 sub, +32) and [128 + 32 sub, +32), lane quarter = warp & 3 (TMEM access rule),
    // sub = warp >> 2 -- loads both (two TMEM loads in flight), says "drained", then publishes the two halves in turn.
    const int e = warp;
    const int quarter = warp & 3;
    const int sub = warp >> 2;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t rx = uint32_t(row_in_tile & 7);
    const uint32_t row_off = uint32_t(row_in_tile >> 3) * 1024u + rx * 128u;
    const uint32_t side_a = smem_u32(side_s);
    uint32_t acc_phase = 0;                  // bit g
    float alpha_s0 = 0.0f, alpha_s1 = 0.0f;  // per-slot alpha partial (layer 7 -> last layer)
    // both CTAs arrive on the LEADER's barriers (the peer with relaxed remote arrives issued after its fences)
    auto arrive = [&](uint64_t* bar) {
      if (leader) mbar_arrive(bar);
      else mbar_arrive_remote(mapa_shared(smem_u32(bar), 0));
    };
    // tile prologue: the slot's accumulator is free (the tile input arrives through the ring)
    if (lane == 0) {
      if (first_tile(0, 0) < n_tiles) arrive(&drained[0]);
      if (first_tile(0, 1) < n_tiles) arrive(&drained[1]);
    }
#pragma unroll 1
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
      const bool active1 = first_tile(iter, 1) < n_tiles;
#pragma unroll 1
      for (int l = 0; l < n_layers; ++l) {
        const MlpLayer& L = prog.layers[l];
        const bool last = (l + 1 == n_layers);
        const bool narrow = (L.n_half == 1);
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          if (g == 1 && !active1) continue;
          const long long t = first_tile(iter, g) + cta_rank;   // may be one past the end in the last pair: rows masked
          const long long grow = t * kTileM + row_in_tile;
          if (e == 0 && lane == 0) ckpt(3, 1000 * int(iter) + 100 * l + 10 * g);
          mbar_wait(&acc_full[g], (acc_phase >> g) & 1u, err_flag, 5);
          acc_phase ^= 1u << g;
          if (e == 0 && lane == 0) ckpt(3, 1000 * int(iter) + 100 * l + 10 * g + 1);
          tc_fence_after();
          if (lane == 0 && (e == 0 || e == EW - 1)) tr(1 + 2 * g + (e != 0), g, l, 3);
          float alpha = (g == 0) ? alpha_s0 : alpha_s1;
          if (l == 0) alpha = 0.0f;
          float rgb[3] = {0.0f, 0.0f, 0.0f};
          const int c0 = sub * 32, c1 = 128 + sub * 32;
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(g * 256);
          uint32_t r0[32], r1[32], p[16];
          if (ADN_SH_DIAG == 0) {
            tmem_ld32(taddr + c0, r0);
            if (!narrow) tmem_ld32(taddr + c1, r1);
            tc_wait_ld();
          }
          // the accumulator lives in registers now: the slot's next layer (or next tile) may overwrite it
          tc_fence_before();
          __syncwarp();
          // (last layer: the warp that reads the other warps' partial sums out of the slot's first hidden block signals after
          // that read -- the next tile's layer-0 epilogue overwrites the block)
          const bool next_layer = !last || first_tile(iter + 1, g) < n_tiles;
          if (lane == 0 && next_layer && !(last && sub == 0)) arrive(&drained[g]);
          if (ADN_SH_DIAG == 0) {
            const int boff = int(L.bias_off);
            const uint32_t st0 = smem_u32(act_ptr(g, int(L.out_blk0) + (c0 >> 6))) + row_off;
            const uint32_t st1 = smem_u32(act_ptr(g, int(L.out_blk0) + (c1 >> 6))) + row_off;
            const uint32_t unit0 = uint32_t(c0 & 63) >> 3;   // same for c1
            if (L.flags & LF_FINAL_RGB) {
              sh_chunk<SK_RGB>(r0, side_a, boff + c0, c0, p, alpha, rgb);
            } else if (!(L.flags & LF_RELU)) {
              sh_chunk<SK_LINEAR>(r0, side_a, boff + c0, c0, p, alpha, rgb);
              sh_store<SK_LINEAR>(p, st0, unit0, rx);
            } else if (L.flags & LF_ALPHA_DOT) {
              sh_chunk<SK_RELU_ALPHA>(r0, side_a, boff + c0, c0, p, alpha, rgb);
              sh_store<SK_RELU_ALPHA>(p, st0, unit0, rx);
            } else {
              sh_chunk<SK_RELU>(r0, side_a, boff + c0, c0, p, alpha, rgb);
              sh_store<SK_RELU>(p, st0, unit0, rx);
            }
            if (L.flags & LF_OUT_ACT) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) arrive(&act_half[2 * g]);
              if (!(L.flags & LF_RELU)) {
                sh_chunk<SK_LINEAR>(r1, side_a, boff + c1, c1, p, alpha, rgb);
              } else if (L.flags & LF_ALPHA_DOT) {
                sh_chunk<SK_RELU_ALPHA>(r1, side_a, boff + c1, c1, p, alpha, rgb);
              } else {
                sh_chunk<SK_RELU>(r1, side_a, boff + c1, c1, p, alpha, rgb);
              }
              sh_store<SK_RELU>(p, st1, unit0, rx);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) arrive(&act_half[2 * g + 1]);
            }
            if (L.flags & LF_FINAL_RGB) {
              // The QW warps of a lane quarter hold partial alpha / rgb sums over their column slices: combine them through
              // shared memory (the slot's hidden blocks are dead once this layer's MMAs have retired).  No barrier after the
              // reads: the next writer of these blocks is the next tile's layer-0 epilogue, whose MMAs cannot start before
              // the reader warps (sub == 0) have arrived on `drained`, which they do after the read.
              float4* scratch = reinterpret_cast<float4*>(act_ptr(g, prog.hid_blk0));
              if (sub > 0) scratch[(sub - 1) * kTileM + row_in_tile] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
              named_bar_sync(12 + quarter, QW * 32);
              if (sub == 0) {
#pragma unroll
                for (int q = 1; q < QW; ++q) {
                  const float4 v = scratch[(q - 1) * kTileM + row_in_tile];
                  rgb[0] += v.x;
                  rgb[1] += v.y;
                  rgb[2] += v.z;
                  alpha += v.w;
                }
                if (grow < rows)
                  reinterpret_cast<float4*>(out)[grow] = make_float4(rgb[0] + side_s[kShRgbB], rgb[1] + side_s[kShRgbB + 1],
                                                                     rgb[2] + side_s[kShRgbB + 2], alpha + side_s[kShAlphaB]);
                __syncwarp();
                if (lane == 0 && next_layer) arrive(&drained[g]);
              }
            }
          } else {
            if (lane == 0 && (L.flags & LF_OUT_ACT)) {
              arrive(&act_half[2 * g]);
              arrive(&act_half[2 * g + 1]);
            }
            if (lane == 0 && last && sub == 0 && next_layer) arrive(&drained[g]);
          }
          if (e == 0 && lane == 0) ckpt(3, 1000 * int(iter) + 100 * l + 10 * g + 2);
          if (lane == 0 && (e == 0 || e == EW - 1)) tr(1 + 2 * g + (e != 0), g, l, 4);
          if (g == 0) alpha_s0 = alpha; else alpha_s1 = alpha;
        }
      }
    }
  }

  if (tracing && threadIdx.x == 0) {
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[7 * 8192 + 2] = clock64();
    trace[7 * 8192 + 3] = gt;
  }
  tc_fence_before();
  cluster_sync_all();   // nobody leaves while its partner may still signal it
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc_cg<2>(tmem_base, 512);
  }
}

cudaError_t set_max_dyn_smem_once(const void* func, int bytes, unsigned long long* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && ((*done >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (dev < 64) *done |= 1ull << dev;   // benign race: the attribute is idempotent
  return cudaSuccess;
}

cudaError_t launch_mlp_sh(const MlpProgram& prog, const uint8_t* wblob, const uint8_t* in_tiles, float* out, const long long* rows_dev,
                          long long rows_host, int* err_flag, int num_sms, cudaStream_t stream, long long* trace) {
  static unsigned long long attr_done = 0;
  const size_t smem = size_t(2) * kShNB * kBlkBytes + size_t(kShStages) * kShStageBytes + size_t(kSideFloats) * 4 + 512 /*barriers*/ +
                      1024 /*alignment slack*/;
  cudaError_t e = set_max_dyn_smem_once(reinterpret_cast<const void*>(mlp_sh_kernel), int(smem), &attr_done);
  if (e != cudaSuccess) return e;
  int grid = (num_sms / 2) * 2;   // persistent: one CTA pair per SM pair
  if (!rows_dev) {
    const long long n_tiles = (rows_host + kTileM - 1) / kTileM;
    const long long need = ((n_tiles + 3) / 4) * 2;   // a pair walks 4 tiles (2 slots x 2 CTAs) per iteration
    if (need < grid) grid = int(need < 2 ? 2 : need);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(unsigned(grid));
  cfg.blockDim = dim3(kShThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, mlp_sh_kernel, prog, wblob, in_tiles, out, rows_dev, rows_host, err_flag, trace);
}

}  // namespace adn
