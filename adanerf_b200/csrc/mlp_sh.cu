// Shading MLP (NeRF 8x256, skip 4, view branch; src/models.py:254-277 of the reference) on tcgen05 / TMEM --
// weight-stationary across two tile slots, N-half pipelined, CTA pairs.
//
// What changed against mlp_umma_kernel<1,2,2> (which this kernel replaces for the shading net) and why:
//   * Tile inputs (position block at layers 0 and 5, view block at the last layer) are A operands that travel through the
//     same FIFO ring as the weights, fetched by the producer right before the step that needs them.  No resident input
//     block per slot: the 32 KB go to the ring (5 stages of 16 KB), and the next tile's input is prefetched for free.
//   * One weight stage feeds BOTH tile slots.  The old kernel streamed every layer once per slot: 24 GB of L2 -> SM
//     traffic per 800x800 frame (32 B / clk / SM while the tensor pipe is busy, 75 % of the measured L2 throughput),
//     and its 4-stage ring covered only 3 x 512 tensor cycles of L2 latency -- the issue loop ran at 82 % of the pipe
//     rate waiting for weights.  Here a stage (two 64-wide K blocks of one 128-column N half, 16 KB per CTA) is used by
//     both slots: half the weight bytes per FLOP and up to 4 x 1024 tensor cycles of prefetch lead.
//   * Every layer is produced as two N halves with their own accumulator columns and barriers (slot g, half h ->
//     TMEM columns 256 g + 128 h).  Half 0's epilogue runs under half 1's MMAs, and the next layer's first K blocks
//     (written by half 0) start while half 1's epilogue is still running: the MMA -> epilogue -> MMA chain that left the
//     tensor pipe idle ~20 % of a layer pair is off the critical path.
//   * Hidden activations are still updated in place (no shared memory left for a second copy), so half 0's epilogue
//     may only STORE once the layer's last MMAs reading blocks out_blk0, out_blk0 + 1 have retired: the issuer
//     commits `lo_free[g]` right after them (MlpLayer::lo_h / lo_s; the K-block order puts those blocks first).
//   * Biases and head vectors are read from a shared-memory copy of the side parameters with warp-uniform 16-byte loads
//     (the register-indexed constant loads of the old epilogue cost 0.5 ms of 5.0), fp32 exact.
//
//   * mbarrier round trips cost a few hundred cycles even when the phase has completed, and a single thread issues one
//     tcgen05.mma per ~50 cycles while the N = 128 MMAs of this kernel carry only 64 tensor cycles each.  A helper warp
//     walking the barriers one after the other, or one issuer warp for both slots, was the kernel's clock.  So there
//     are two issuer warps, one per slot, each waiting for everything a step needs at once, one mbarrier per lane
//     (weight stage, input stage, the slot's accumulator half drained / activation blocks written), with an early
//     non-blocking probe of the next weight stage.  Both CTAs of a pair arrive on the LEADER's barriers: the peer's
//     epilogue warps with relaxed remote arrives after their proxy fence, its two otherwise idle issuer warps forward
//     "my share of the stage has landed" -- one barrier per dependency instead of a local + forwarded pair.
//   * Epilogue warps are split by slot (8 per slot, 64 columns = one full swizzled row of one block per warp and event)
//     and keep two TMEM loads in flight: the per-chunk latency chain (TMEM read, pack, store, proxy fence), not the
//     instruction count, bounds the epilogue.
//
// What it buys and what bounds it now (profiles/trace_sh.py): 3-12 % over the round-1 kernel in
// same-box A/B runs (the power-capped boxes gain most), half the L2 -> SM weight traffic, fp32-exact biases; 60 % tensor pipe
// active under ncu.  A tile pair (580 MMAs, 37.4 k tensor cycles at the 64-cycle rate) takes ~65 k cycles.  Timing probes
// with deliberately incomplete kernels (ADN_SH_DIAG builds, wrong results, kept in git history) split the gap three ways:
//     47.2 k  empty epilogue, no weight / input traffic at all: the two issuer warps alone.  A step is ~200 executed
//             instructions (decode of the schedule word, ring bookkeeping, lane-parallel barrier wait, 8 MMAs at ~50 cycles
//             each, commits); the tensor queue is ~2 MMAs deep, so the pipe idles whenever both issuers are between steps;
//     53.0 k  + the weight ring: every step still waits ~0.3 k cycles for its stage although the copy was issued ~4.5 k cycles
//             earlier (5 stages of 16 KB cannot cover the fill latency when two are borrowed by tile inputs);
//     59.0 k  real epilogue, no weight traffic;      65 k  everything.
// Ruled out by probes: the shared-memory A operand (A read from TMEM instead: 53.7 k against 53.0 k, no change -- the 91
// cycles per N = 128 MMA are not an operand-bandwidth limit); L2 hot-spotting on the weight lines (1 / 4 / 16 / 37 replicas
// of the blob, MlpProgram::w_copies: 4.89-4.96 ms, no change).  Measured and dropped (git history): tile inputs fetched
// once per N half instead of once per layer (frees two of the five ring stages for four steps: +2 % frames/s in a same-box A/B,
// but the determinism test saw frame-to-frame differences; reverted.  Cause, found afterwards on paper: an issuer commits the
// release of the OTHER slot's input stage (`in_other`) without ever having observed that stage's `w_full`, so its arrival for
// pass n + 1 of that ring stage can reach `w_empty` before the other issuer's arrival for pass n -- the phase then completes
// on two arrivals of the same warp and the producer refills a stage the other slot's MMAs still read.  With the release in
// the fetch step this needs a lag of less than one step between the issuers; in the shipped schedule only the last layer's
// view block is released in its fetch step and the lag needed is three steps (> 3000 cycles, never seen: 0 differing frames
// in profiles/determinism_stress.py).  Fix to apply with a GPU at hand: at fetch steps let a fifth lane of each leader issuer
// wait for `w_full[in_other]` as well before the step's commits); an issue loop
// specialised per schedule word (compile-time stages, 65 KB of code: 5.87 ms against 4.88 -- the footprint costs more
// instruction-cache misses than the decode saved); the issuing warp software-pipelined (next step's schedule word, ring stages
// and one non-blocking probe per barrier between the MMAs of the current step: the waits shrink from ~550 to ~400 cycles but
// the step grows from 1350 to 2170 -- anything placed between the MMAs of a step starves the tensor queue: 6.55 ms); a per-layer specialised epilogue with immediate bias operands (160 KB:
// 2x slower, same reason); N = 256 MMAs with alternating slots, per-slot weight passes, an early "accumulator drained"
// signal and half-wise activation hand-off (5.3-5.7 ms).
//
// Roles (19 warps): 16 epilogue warps, weight producer, 2 MMA issuers (leader CTA) / barrier forwarders (peer CTA).
// Schedule walked by every role:
//   for tile group:  for layer:  for half h:  for stage s (K blocks 2s, 2s+1):  for slot g:  8 MMAs (M 256, N 128, K 16)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "mlp_umma.cuh"
#include "ptx.cuh"

namespace adn {

namespace {

#ifndef ADN_SH_DIAG
// Timing experiments only (wrong results): 1 = epilogue events do nothing, 2 = epilogue without shared-memory stores.
#define ADN_SH_DIAG 0
#endif
constexpr int kShThreads = 608;            // 19 warps
constexpr int kShStages = 5;               // ring stages per CTA (a weight stage feeds both slots: 1024 tensor cycles)
constexpr int kShStageBytes = kBlkBytes;   // weights: two K blocks x this CTA's 64 rows of the 128-row N half; or one input block
constexpr int kShHalfBlk = kBlkBytes / 2;  // one K block of weights, this CTA's 64 B rows
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

// One epilogue event of one warp: accumulator columns [c, c + 64) of one (slot, half) for this thread's row = one full
// 128-byte swizzled row of one activation block.  The second TMEM load is in flight while the first chunk is processed.
//   taddr : TMEM address of (lane quarter, slot, column c);  st_row : shared-memory address of this row in that block
template <int KIND>
__device__ __forceinline__ void sh_event(uint32_t side_s, int boff, int c, uint32_t taddr, uint32_t st_row, uint32_t rx, bool wait_lo,
                                         uint64_t* lo_bar, uint32_t lo_parity, int* err_flag, float& alpha, float (&rgb)[3]) {
  if (ADN_SH_DIAG == 1) {
    if (KIND != SK_RGB && wait_lo) mbar_wait(lo_bar, lo_parity, err_flag, 6);
    return;
  }
  uint32_t r0[32], r1[32], p[16];
  tmem_ld32(taddr, r0);
  tc_wait_ld();
  tmem_ld32(taddr + 32, r1);
  sh_chunk<KIND>(r0, side_s, boff, c, p, alpha, rgb);
  if (KIND != SK_RGB) {
    // in-place hazard: this layer's half-1 MMAs may still be reading the block written below
    if (wait_lo) mbar_wait(lo_bar, lo_parity, err_flag, 6);
    if (ADN_SH_DIAG == 2) {
      if (p[0] == 0x12345u && p[7] == 0x777u) st_shared_v4(st_row, p[0], p[1], p[2], p[3]);   // keeps the math alive
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_shared_v4(st_row + ((uint32_t(q) ^ rx) << 4), p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    }
  }
  tc_wait_ld();
  sh_chunk<KIND>(r1, side_s, boff + 32, c + 32, p, alpha, rgb);
  if (KIND != SK_RGB) {
    if (ADN_SH_DIAG == 2) {
      if (p[0] == 0x12345u && p[7] == 0x777u) st_shared_v4(st_row, p[0], p[1], p[2], p[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) st_shared_v4(st_row + ((uint32_t(4 + q) ^ rx) << 4), p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(kShThreads, 1)
mlp_sh_kernel(const __grid_constant__ MlpProgram prog, const uint8_t* __restrict__ wblob, const uint8_t* __restrict__ in_tiles,
              float* __restrict__ out, const long long* __restrict__ rows_dev, long long rows_host, int* err_flag, long long* trace) {
  constexpr int STAGES = kShStages, STAGE_BYTES = kShStageBytes, HALF = kShHalfBlk, NB = kShNB;
  constexpr int GW = 8;   // epilogue warps per slot
  constexpr int kProducerWarp = 16, kMmaWarp = 17;   // issuers / forwarders: warps 17 (slot 0), 18 (slot 1)
  // named barriers 1, 2: rgb exchange of the slot's epilogue warps

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* act = smem;                                        // [2 slots][NB] hidden blocks
  uint8_t* ring = act + size_t(2) * NB * kBlkBytes;           // [STAGES]
  float* side_s = reinterpret_cast<float*>(ring + size_t(STAGES) * STAGE_BYTES);   // copy of prog.side (biases, head vectors)
  uint64_t* bars = reinterpret_cast<uint64_t*>(side_s + kSideFloats);
  uint64_t* w_full = bars;                   // [STAGES] this CTA's share of the stage has landed (leader: and the peer's)
  uint64_t* w_empty = w_full + STAGES;       // [STAGES] the MMAs reading the stage have retired (both issuers, both CTAs)
  uint64_t* acc_full = w_empty + STAGES;     // [2 g + h] accumulator half complete
  uint64_t* act_ready = acc_full + 4;        // [2 g + h] leader only: both CTAs' epilogues of that half are done (TMEM drained, A blocks written)
  uint64_t* lo_free = act_ready + 4;         // [g] the blocks half 0's epilogue overwrites are no longer read by this layer's MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lo_free + 2);

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
      mbar_init(&w_empty[s], 2);               // one commit per issuer warp
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&act_ready[i], 2 * GW);        // leader only: the slot's epilogue warps of BOTH CTAs arrive here
    }
    mbar_init(&lo_free[0], 1);
    mbar_init(&lo_free[1], 1);
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
  // word 0 = count; code = slot << 16 | (2 layer + half) << 8 | event.
  const bool tracing = (trace != nullptr) && (blockIdx.x == 0);
  int tr_n = 0;
  auto tr = [&](int region, int g, int lh, int ev) {
    if (tracing && tr_n < 4000) {
      long long* base = trace + region * 8192;
      base[2 + 2 * tr_n] = clock64();
      base[3 + 2 * tr_n] = (long long)((g << 16) | (lh << 8) | ev);
      base[0] = ++tr_n;
    }
  };
  if (tracing && threadIdx.x == 0) {   // region 7: kernel start / end in SM clocks and in ns (the real SM frequency)
    long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[7 * 8192 + 0] = clock64();
    trace[7 * 8192 + 1] = gt;
  }

  if (warp == kProducerWarp) {
    // ===================================================================== producer (weights and tile inputs)
    // Weight stage of step (l, h, s): K blocks 2s, 2s+1 of N half h, this CTA's rows [64 rank, +64) of each [128 x 64]
    // tile.  A step flagged "fetch input" is followed by one stage per slot holding that slot's tile input block.
    if (lane == 0) {
      const uint8_t* wcopy = wblob + size_t((blockIdx.x >> 1) % prog.w_copies) * prog.w_stride;
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
        for (int l = 0; l < n_layers; ++l) {
          const MlpLayer& L = prog.layers[l];
          const int n_kb = L.n_kb, n_steps = prog.sh_steps[l];
          for (int i = 0; i < n_steps; ++i) {
            const uint32_t w = prog.sh_sched[l][i];
            const int h = int((w >> 14) & 1u), s = int((w >> 20) & 3u);
            const uint8_t* src = wcopy + L.w_off + (size_t(h) * n_kb + 2 * s) * kBlkBytes + cta_rank * HALF;
            tr(5, stage, 2 * l + h, 8);
            mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 1);
            tr(5, stage, 2 * l + h, 9);
            uint8_t* dst = ring + size_t(stage) * STAGE_BYTES;
            const int nkb = (w & (1u << 8)) ? 2 : 1;
            mbar_arrive_expect_tx(&w_full[stage], uint32_t(nkb) * HALF);
            bulk_g2s(dst, src, HALF, &w_full[stage]);
            if (nkb == 2) bulk_g2s(dst + HALF, src + kBlkBytes, HALF, &w_full[stage]);
            advance();
            if (w & (1u << 17)) {   // this step's K block is the tile input: one stage per slot
              const uint32_t off = (w & (1u << 19)) ? prog.in1_off : prog.in0_off;
              for (int g = 0; g < 2; ++g) {
                mbar_wait(&w_empty[stage], phase ^ 1, err_flag, 2);
                const long long t = first_tile(iter, g) + cta_rank;
                if (t < n_tiles) {
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
  } else if (warp >= kMmaWarp) {
    // ================================================== MMA issuers (leader CTA) / barrier forwarders (peer CTA)
    // Warp kMmaWarp + g serves slot g.  Per step it needs: the weight stage (lanes 6, 7), its slot's input stage when the
    // step fetches one (lanes 4, 5), and -- the first time in a layer -- the previous layer's half-0 / half-1 epilogue of
    // its slot (lanes 0, 1 / 2, 3); even lanes watch this CTA's barrier, odd lanes (leader only) the peer's forwarded
    // copy.  All of them wait at once.  Leader: then one elected lane issues the step's MMAs and commits.  Peer: the even
    // lanes forward with relaxed remote arrives (they have acquired the local barrier first; a release at cluster scope
    // would cost a MEMBAR.ALL.GPU); the weight stage is forwarded by the slot-0 warp only.
    // Both leader warps walk every step -- also when their slot has no tile in the last tile group -- and both commit
    // w_empty of every stage a step releases (count 2).
    const int g = warp - kMmaWarp;
    constexpr uint32_t idesc128 = make_idesc_bf16(256, 128);
    const uint64_t desc_hi = make_desc_sw128(0) & 0xFFFFFFFF00000000ull;
    const uint32_t desc_lo_const = uint32_t(make_desc_sw128(0) & 0xFFFF0000ull);
    auto lo_of = [&](uint32_t addr) -> uint32_t { return desc_lo_const | (addr >> 4); };
    const uint32_t ring_lo = lo_of(smem_u32(ring));
    const uint32_t act_lo0 = lo_of(smem_u32(act_ptr(g, 1)));
    auto desc = [&](uint32_t lo) -> uint64_t { return desc_hi | uint64_t(lo); };
    int stage = 0, in_stage = 0, in_other = 0;
    uint32_t phase = 0, ar_phase = 0;   // ar_phase: bit hh = parity of act_ready[2 g + hh]
    bool w_seen = false;                // lane 3: the early probe has already seen this step's weight stage complete
    auto advance = [&]() {
      if (++stage == STAGES) {
        stage = 0;
        phase ^= 1;
      }
    };
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, 0) >= n_tiles) break;
      const bool active = first_tile(iter, g) < n_tiles;
      for (int l = 0; l < n_layers; ++l) {
        const int n_steps = prog.sh_steps[l];
        uint32_t seen = 0;
        for (int i = 0; i < n_steps; ++i) {
          const uint32_t w = prog.sh_sched[l][i];
          const uint32_t need = (active && leader) ? ((w >> 9) & 3u) : 0u;
          seen |= need;
          const int w_stage = stage;
          const uint32_t w_par = phase;
          advance();
          uint32_t in_par = 0;
          const bool fetch = (w & (1u << 17)) != 0;
          if (fetch) {   // input stages: slot 0's, then slot 1's
            if (g == 1) {
              in_other = stage;
              advance();
            }
            in_stage = stage;
            in_par = phase;
            advance();
            if (g == 0) {
              in_other = stage;
              advance();
            }
          }
          if (lane == 0 && g == 0) tr(0, i, 2 * l + int((w >> 14) & 1u), 0);
          {
            // one barrier per lane: 0, 1: previous layer's half-0 / half-1 epilogue of this slot (both CTAs arrive on the
            // leader's barrier); 2: this slot's input stage; 3: the weight stage (in the leader: own share + the peer's forward)
            const bool mine = lane < 2 ? ((need >> lane) & 1u) != 0 : (lane == 2 ? fetch : (lane == 3 && !w_seen));   // everybody waits for the weight
            // stage: a warp that ran ahead by a full ring phase would alias the parity of a later phase
            uint64_t* bar = lane < 2 ? &act_ready[2 * g + (lane & 1)] : (lane == 2 ? &w_full[in_stage] : &w_full[w_stage]);
            const uint32_t parity = lane < 2 ? ((ar_phase >> (lane & 1)) & 1u) : (lane == 2 ? in_par : w_par);
            mbar_wait_lanes(bar, parity, mine, err_flag, 3);
            if (!leader && ((lane == 2 && fetch) || (lane == 3 && g == 0))) mbar_arrive_remote(mapa_shared(smem_u32(bar), 0));
            __syncwarp();
            ar_phase ^= need;
          }
          // Early, non-blocking probe of the NEXT step's weight stage (ring position `stage` after the advances above): its
          // round trip overlaps the MMA issue below instead of draining the tensor pipe's short queue at the next step.
          w_seen = (lane == 3) ? mbar_test(&w_full[stage], phase) : false;
          if (leader) {
            tc_fence_after();
            if (lane == 0 && g == 0) tr(0, g, 2 * l + int((w >> 14) & 1u), (w & (1u << 15)) ? 1 : 6);
            const uint32_t h = (w >> 14) & 1u;
            const uint32_t acc = ((w >> 15) & 1u) ^ 1u;   // first step of the half: overwrite
            const uint32_t b = ring_lo + uint32_t(w_stage) * (STAGE_BYTES >> 4);
            const bool input = (w & (1u << 16)) != 0;
            const uint32_t a0 = input ? ring_lo + uint32_t(in_stage) * (STAGE_BYTES >> 4) : act_lo0 + ((w & 15u) - 1u) * (kBlkBytes >> 4);
            const uint32_t a1 = act_lo0 + (((w >> 4) & 15u) - 1u) * (kBlkBytes >> 4);
            const uint32_t d = tmem_base + uint32_t(g * 256) + h * 128u;
            if (elect_one()) {
              if (active) {
                umma_bf16_cg<2>(d, desc(a0), desc(b), idesc128, acc);
                umma_bf16_cg<2>(d, desc(a0 + 2), desc(b + 2), idesc128, 1u);
                if (!(w & (1u << 19))) {   // the view block carries 27 features: two K steps
                  umma_bf16_cg<2>(d, desc(a0 + 4), desc(b + 4), idesc128, 1u);
                  umma_bf16_cg<2>(d, desc(a0 + 6), desc(b + 6), idesc128, 1u);
                }
                if (w & (1u << 8)) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) umma_bf16_cg<2>(d, desc(a1 + 2 * k), desc(b + (HALF >> 4) + 2 * k), idesc128, 1u);
                }
                if (w & (1u << 12)) umma_commit_cg<2>(&acc_full[2 * g + int(h)]);
                if (w & (1u << 13)) umma_commit_cg<2>(&lo_free[g]);
              }
              umma_commit_cg<2>(&w_empty[w_stage]);
              if (w & (1u << 18)) {   // the tile-input stages of both slots are released here (every stage expects two commits)
                umma_commit_cg<2>(&w_empty[in_stage]);
                umma_commit_cg<2>(&w_empty[in_other]);
              }
            }
            __syncwarp();
            if (lane == 0 && g == 0 && (w & (1u << 12))) tr(0, g, 2 * l + int(h), 2);
          }
        }
        // one completion per layer and (slot, half): consume what this layer did not need (leader only; not reached by the
        // NeRF program, every layer of which needs both halves)
        const uint32_t rest = (active && leader) ? (3u & ~seen) : 0u;
        if (rest) {
          mbar_wait_lanes(&act_ready[2 * g + (lane & 1)], (ar_phase >> (lane & 1)) & 1u, lane < 2 && ((rest >> lane) & 1u) != 0, err_flag, 9);
          __syncwarp();
          ar_phase ^= rest;
        }
      }
    }
  } else {
    // ============================================================================ epilogue (2 x 8 warps)
    // Warps 0..7 serve slot 0, warps 8..15 slot 1; events in the issuer's completion order: for layer, for half.  Every warp
    // takes 64 columns of the [128 x 128] accumulator half: lane quarter = warp & 3 (TMEM access rule), column half =
    // (warp >> 2) & 1 -- one full swizzled row of one activation block.
    const int g = warp >> 3;
    const int e = warp & 7;
    const int quarter = warp & 3;
    const int sub = (warp >> 2) & 1;
    const int row_in_tile = quarter * 32 + lane;
    const uint32_t rx = uint32_t(row_in_tile & 7);
    const uint32_t row_off = uint32_t(row_in_tile >> 3) * 1024u + rx * 128u;
    const uint32_t side_a = smem_u32(side_s);
    uint32_t acc_phase = 0, lo_phase = 0;   // bit h / single bit
    float alpha = 0.0f;                     // alpha partial (layer 7 -> last layer)
    // "this warp's share of (slot, half) is drained and written": both CTAs arrive on the LEADER's barrier (the peer with a
    // relaxed remote arrive issued after its proxy fence; a release at cluster scope would cost a MEMBAR.ALL.GPU)
    auto arrive_act = [&](int bit) {
      if (leader) mbar_arrive(&act_ready[bit]);
      else mbar_arrive_remote(mapa_shared(smem_u32(&act_ready[bit]), 0));
    };
    // tile prologue: hand the (free) accumulators / activation blocks to layer 0 (the tile input arrives through the ring)
    auto tile_prologue = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        arrive_act(2 * g);
        arrive_act(2 * g + 1);
      }
    };
    if (first_tile(0, g) < n_tiles) tile_prologue();
#pragma unroll 1
    for (long long iter = 0;; ++iter) {
      if (first_tile(iter, g) >= n_tiles) break;
      const long long t = first_tile(iter, g) + cta_rank;   // may be one past the end in the last pair: rows masked
      const long long grow = t * kTileM + row_in_tile;
#pragma unroll 1
      for (int l = 0; l < n_layers; ++l) {
        const MlpLayer& L = prog.layers[l];
        const bool last = (l + 1 == n_layers);
        if (l == 0) alpha = 0.0f;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const int bit = 2 * g + h;
          if (h >= L.n_half) {   // single-half layer: nothing to drain, the barrier still completes once per layer
            if (!last) {
              __syncwarp();
              if (lane == 0) arrive_act(bit);
            }
            continue;
          }
#ifndef ADN_SPIN_EPI
#define ADN_SPIN_EPI 0   // 1: epilogue warps poll acc_full / lo_free with test_wait (no hardware suspend) -- experiment
#endif
          if (ADN_SPIN_EPI) mbar_spin(&acc_full[bit], (acc_phase >> h) & 1u, err_flag, 5);
          else mbar_wait(&acc_full[bit], (acc_phase >> h) & 1u, err_flag, 5);
          acc_phase ^= 1u << h;
          tc_fence_after();
          if (lane == 0 && e == 0) tr(1 + g, g, 2 * l + h, 3);
          float rgb[3] = {0.0f, 0.0f, 0.0f};
          const int c = h * 128 + sub * 64;   // this warp's accumulator columns [c, c + 64)
          const uint32_t taddr = tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(g * 256 + c);
          const uint32_t st_row = smem_u32(act_ptr(g, int(L.out_blk0) + (c >> 6))) + row_off;
          const bool wait_lo = (h == 0) && (L.lo_h != 0xFF);
          const int boff = int(L.bias_off) + c;
          if (L.flags & LF_FINAL_RGB) sh_event<SK_RGB>(side_a, boff, c, taddr, st_row, rx, false, &lo_free[g], lo_phase, err_flag, alpha, rgb);
          else if (!(L.flags & LF_RELU)) sh_event<SK_LINEAR>(side_a, boff, c, taddr, st_row, rx, wait_lo, &lo_free[g], lo_phase, err_flag, alpha, rgb);
          else if (L.flags & LF_ALPHA_DOT) sh_event<SK_RELU_ALPHA>(side_a, boff, c, taddr, st_row, rx, wait_lo, &lo_free[g], lo_phase, err_flag, alpha, rgb);
          else sh_event<SK_RELU>(side_a, boff, c, taddr, st_row, rx, wait_lo, &lo_free[g], lo_phase, err_flag, alpha, rgb);
          if (wait_lo) lo_phase ^= 1u;   // consumed inside the event
          if (L.flags & LF_FINAL_RGB) {
            // The two warps of a lane quarter hold partial alpha / rgb sums over their 64 columns: combine them through shared
            // memory (the slot's hidden blocks are dead once this layer's MMAs have retired).  No barrier after the reads:
            // the next writer of these blocks is the next tile's layer-0 epilogue, which cannot start before all of the
            // slot's epilogue warps -- the readers included -- have arrived on act_ready.
            float4* scratch = reinterpret_cast<float4*>(act_ptr(g, prog.hid_blk0));
            if (sub == 1) scratch[row_in_tile] = make_float4(rgb[0], rgb[1], rgb[2], alpha);
            named_bar_sync(1 + g, GW * 32);
            if (sub == 0) {
              const float4 q = scratch[row_in_tile];
              if (grow < rows)
                reinterpret_cast<float4*>(out)[grow] = make_float4(rgb[0] + q.x + side_s[kShRgbB], rgb[1] + q.y + side_s[kShRgbB + 1],
                                                                   rgb[2] + q.z + side_s[kShRgbB + 2], alpha + q.w + side_s[kShAlphaB]);
            }
          }
          if ((L.flags & LF_OUT_ACT) && (ADN_SH_DIAG == 0 || ADN_SH_DIAG == 2)) fence_proxy_async_smem();
          if (lane == 0 && e == 0) tr(1 + g, g, 2 * l + h, 4);
          if (!last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_act(bit);
          } else if (h == L.n_half - 1) {
            if (first_tile(iter + 1, g) < n_tiles) tile_prologue();   // the slot's next tile starts right away
          }
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
  if (dev < 64 && ((__atomic_load_n(done, __ATOMIC_ACQUIRE) >> dev) & 1ull)) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  if (dev < 64) __atomic_fetch_or(done, 1ull << dev, __ATOMIC_RELEASE);   // two threads may both set the (idempotent) attribute
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
