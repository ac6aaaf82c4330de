// SIMT stages (see stages.cuh).  All position math uses explicit round-to-nearest intrinsics in the
// reference's operation order so results track the PyTorch fp32 path (no FMA contraction where torch
// has separate mul/add; an FMA chain where ATen's K=3 bmm uses one).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "mlp_umma.cuh"
#include "posenc.cuh"
#include "ptx.cuh"
#include "stages.cuh"

namespace adn {

constexpr int kNFreqPos = 10;
constexpr int kNFreqDir = 4;
constexpr int kFeat = 90;

// ------------------------------------------------------------------------------------ stage 0a
__device__ __forceinline__ void pixel_dir(const CameraRays& cam, long long ray, float (&d)[3]) {
  const int y = cam.row0 + int(ray / cam.W);
  const int x = int(ray % cam.W);
  const double rx = __dadd_rn(cam.start_x, __dmul_rn(cam.x_pp, double(x)));
  const double ry = __dadd_rn(cam.start_y, __dmul_rn(cam.y_pp, double(y)));
  const double rz = cam.focal;
  const double n = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)));
  d[0] = float(__ddiv_rn(rx, n));
  d[1] = float(-__ddiv_rn(ry, n));
  d[2] = float(-__ddiv_rn(rz, n));
}

__global__ void gen_dirs_kernel(const __grid_constant__ CameraRays cam, long long n_rays, float* __restrict__ dirs) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n_rays) return;
  float d[3];
  pixel_dir(cam, i, d);
  dirs[3 * i + 0] = d[0];
  dirs[3 * i + 1] = d[1];
  dirs[3 * i + 2] = d[2];
}

cudaError_t launch_gen_dirs(const CameraRays& cam, long long n_rays, float* d_dirs, cudaStream_t s) {
  if (n_rays <= 0) return cudaSuccess;
  gen_dirs_kernel<<<unsigned((n_rays + 255) / 256), 256, 0, s>>>(cam, n_rays, d_dirs);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------ stage 0b
// SpherePosDir.batch (src/features.py:845-899): one thread per ray.
template <bool FROM_CAMERA, int NFD = kNFreqDir, int NFP = kNFreqPos>
__global__ void __launch_bounds__(128)
stage0_kernel(const __grid_constant__ SceneDev sc, const __grid_constant__ PoseDev pd, const float* __restrict__ dirs,
              const __grid_constant__ CameraRays cam, long long n_rays, float* __restrict__ x0, float* __restrict__ ray_o,
              float* __restrict__ ray_d, uint8_t* __restrict__ tiles0) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // grid covers whole 128-ray tiles
  float f[128];
#pragma unroll
  for (int j = 0; j < 128; ++j) f[j] = 0.0f;
  if (i < n_rays) {
    float d[3];
    if (FROM_CAMERA) {
      pixel_dir(cam, i, d);
    } else {
      d[0] = dirs[3 * i + 0];
      d[1] = dirs[3 * i + 1];
      d[2] = dirs[3 * i + 2];
    }
    // nds = R * d : ATen bmm with K = 3 accumulates as an FMA chain over k = 0,1,2 (:858-859)
    float nds[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      nds[r] = __fmaf_rn(pd.rot[3 * r + 2], d[2], __fmaf_rn(pd.rot[3 * r + 1], d[1], __fmul_rn(pd.rot[3 * r + 0], d[0])));
    // compute_ray_offset (:769-791)
    float omc[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) omc[a] = __fsub_rn(pd.pose[a], sc.c[a]);
    const float udot = __fadd_rn(__fadd_rn(__fmul_rn(omc[0], nds[0]), __fmul_rn(omc[1], nds[1])), __fmul_rn(omc[2], nds[2]));
    const float omc2 = __fadd_rn(__fadd_rn(__fmul_rn(omc[0], omc[0]), __fmul_rn(omc[1], omc[1])), __fmul_rn(omc[2], omc[2]));
    const float delta = __fsub_rn(__fmul_rn(udot, udot), __fsub_rn(omc2, sc.r2));
    const float t = __fadd_rn(-udot, __fsqrt_rn(fmaxf(delta, 0.0f)));
    float p[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) p[a] = __fadd_rn(pd.pose[a], __fmul_rn(nds[a], t));
    const float nn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(nds[0], nds[0]), __fmul_rn(nds[1], nds[1])), __fmul_rn(nds[2], nds[2])));
    float dn[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) dn[a] = __fdiv_rn(nds[a], nn);
    constexpr int F0 = 6 + 6 * (NFD + NFP);       // 90 ("10-4") or 30 ("2-2")
    posenc3<NFD>(dn, f);                           // 27 / 15: direction block FIRST (:868)
    posenc3<NFP>(p, f + 3 + 6 * NFD);              // 63 / 15
    if (ray_o) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ray_o[3 * i + a] = p[a];
        ray_d[3 * i + a] = nds[a];
      }
    }
    if (x0) {
#pragma unroll
      for (int j = 0; j < F0; ++j) x0[i * F0 + j] = f[j];
    }
  }
  if (tiles0) {
    // packed MLP0 input: per tile [hi blk0 | hi blk1 | lo blk0 | lo blk1], 16 KB each.  The block (= one
    // 128-ray tile) assembles the 64 KB image in shared memory and one thread hands it to the TMA engine
    // (bulk shared -> global copy): no strided 16-byte global stores.
    extern __shared__ __align__(1024) uint8_t s_tile[];
    const long long t = i >> 7;
    const uint32_t r = uint32_t(i & 127);
#pragma unroll
    for (int b = 0; b < 2; ++b) {
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const float* v = f + b * 64 + ch * 8;
        uint4 hi;
        hi.x = bf16x2(v[0], v[1]);
        hi.y = bf16x2(v[2], v[3]);
        hi.z = bf16x2(v[4], v[5]);
        hi.w = bf16x2(v[6], v[7]);
        const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w};
        uint32_t lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float l0 = v[2 * e + 0] - __uint_as_float(hw[e] << 16);
          const float l1 = v[2 * e + 1] - __uint_as_float(hw[e] & 0xFFFF0000u);
          lw[e] = bf16x2(l0, l1);
        }
        const uint32_t off = sw128_offset(r, uint32_t(ch * 8));
        *reinterpret_cast<uint4*>(s_tile + b * kBlkBytes + off) = hi;
        *reinterpret_cast<uint4*>(s_tile + (2 + b) * kBlkBytes + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_s2g(tiles0 + size_t(t) * (4 * kBlkBytes), s_tile, 4 * kBlkBytes);
      bulk_commit();
      bulk_wait_all();
    }
  }
}

cudaError_t launch_stage0(const SceneDev& sc, const PoseDev& pd, const float* d_dirs, const CameraRays* cam,
                          long long n_rays, float* d_x0, float* d_ray_o, float* d_ray_d, uint8_t* d_tiles0,
                          cudaStream_t s) {
  if (n_rays <= 0) return cudaSuccess;
  const long long n_pad = ((n_rays + kTileM - 1) / kTileM) * kTileM;
  const unsigned grid = unsigned((n_pad + 127) / 128);
  CameraRays c{};
  const size_t smem = d_tiles0 ? size_t(4 * kBlkBytes) : 0;
  static unsigned long long attr_cam = 0, attr_rays = 0;   // per device
  if (set_max_dyn_smem_once(reinterpret_cast<const void*>(stage0_kernel<true>), 4 * kBlkBytes, &attr_cam) != cudaSuccess ||
      set_max_dyn_smem_once(reinterpret_cast<const void*>(stage0_kernel<false>), 4 * kBlkBytes, &attr_rays) != cudaSuccess)
    return cudaGetLastError();
  if (cam) c = *cam;
  if (sc.n_freq_pos0 == 2 && sc.n_freq_dir0 == 2) {   // "2-2" (NDC configs): 30 features, same tile image (columns 30.. are zero)
    static unsigned long long attr_cam22 = 0, attr_rays22 = 0;
    if (set_max_dyn_smem_once(reinterpret_cast<const void*>(stage0_kernel<true, 2, 2>), 4 * kBlkBytes, &attr_cam22) != cudaSuccess ||
        set_max_dyn_smem_once(reinterpret_cast<const void*>(stage0_kernel<false, 2, 2>), 4 * kBlkBytes, &attr_rays22) != cudaSuccess)
      return cudaGetLastError();
    if (cam) stage0_kernel<true, 2, 2><<<grid, 128, smem, s>>>(sc, pd, d_dirs, c, n_rays, d_x0, d_ray_o, d_ray_d, d_tiles0);
    else stage0_kernel<false, 2, 2><<<grid, 128, smem, s>>>(sc, pd, d_dirs, c, n_rays, d_x0, d_ray_o, d_ray_d, d_tiles0);
  } else if (cam) {
    stage0_kernel<true><<<grid, 128, smem, s>>>(sc, pd, d_dirs, c, n_rays, d_x0, d_ray_o, d_ray_d, d_tiles0);
  } else {
    stage0_kernel<false><<<grid, 128, smem, s>>>(sc, pd, d_dirs, c, n_rays, d_x0, d_ray_o, d_ray_d, d_tiles0);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- stage 2
// FromClassifiedDepthAdaptive.generate (src/nerf_raymarch_common.py:699-757) + mask compaction
// (src/features.py:445-446,481-484), single pass:
//   * one warp per ray, lane l owns cells 4l..4l+3 (one coalesced 512-byte row load),
//   * survivors = cells >= thr; more than K survivors -> K rounds of warp arg-max (value descending,
//     ties lower cell first); none -> the arg-max cell; order inside a ray = ascending cell = ascending z,
//   * per-CTA exclusive scan of the counts + decoupled look-back across CTAs (dynamic tickets), so the
//     packed order is deterministic ray-major (the torch boolean-mask order) with no host sync.
constexpr int kS2Rays = 64;     // rays per CTA
constexpr int kS2Threads = 256;  // 8 warps x 8 rays

// Order-preserving map float -> uint32 (larger float <-> larger key); -0.0 is folded onto +0.0 so it ties
// with it like a float compare does.  Every real input maps to a key >= 0x007FFFFF, so 0 means "no entry".
__device__ __forceinline__ uint32_t order_key(float v) {
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ void cswap_desc(unsigned long long& a, unsigned long long& b) {
  const unsigned long long hi = a > b ? a : b, lo = a > b ? b : a;
  a = hi;
  b = lo;
}

// Returns the 4 selection masks (bit l of sel[j] <-> cell 4l+j) and the count, warp uniform.
// More than K survivors: K rounds of "pop the best head".  Each lane keeps its (up to 4) candidates as
// 64-bit composites (order_key << 2 | 3 - j), sorted descending, so the head of every lane is its best
// remaining cell; one REDUX.MAX over the heads + a ballot finds the winner (ties: lowest lane = lowest
// cell, and inside a lane the lower j sorts first), and only the winning lane pops.
__device__ __forceinline__ int select_cells(const float4 v4, float thr, int K, int lane, uint32_t (&sel)[4]) {
  const float v[4] = {v4.x, v4.y, v4.z, v4.w};
  uint32_t act[4];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    act[j] = __ballot_sync(0xffffffffu, v[j] >= thr);
    cnt += __popc(act[j]);
  }
  if (cnt > 0 && cnt <= K) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sel[j] = act[j];
    return cnt;
  }
  const bool fallback = (cnt == 0);          // nothing >= thr: the arg-max cell (:748-749)
  const int need = fallback ? 1 : K;
  unsigned long long h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool cand = fallback || ((act[j] >> lane) & 1u);
    h[j] = cand ? (((unsigned long long)order_key(v[j]) << 2) | (unsigned long long)(3 - j)) : 0ull;
  }
  cswap_desc(h[0], h[1]);
  cswap_desc(h[2], h[3]);
  cswap_desc(h[0], h[2]);
  cswap_desc(h[1], h[3]);
  cswap_desc(h[1], h[2]);
  uint32_t mine = 0;
  for (int round = 0; round < need; ++round) {
    const uint32_t head = uint32_t(h[0] >> 2);
    const uint32_t m = __reduce_max_sync(0xffffffffu, head);
    const uint32_t who = __ballot_sync(0xffffffffu, head == m);
    if (lane == __ffs(who) - 1) {
      mine |= 1u << (3 - int(h[0] & 3ull));
      h[0] = h[1];
      h[1] = h[2];
      h[2] = h[3];
      h[3] = 0ull;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) sel[j] = __ballot_sync(0xffffffffu, (mine >> j) & 1u);
  return need;
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Tile bookkeeping, executed by one full warp.  Tiles are numbered by dynamic tickets, so every predecessor a tile
// spins on has already started.  A tile's state word carries flag and value together (no fence needed).
// State word: [63:62] flag (1 = tile aggregate, 2 = inclusive prefix), [61:40] launch epoch, [39:0] value.  A word whose
// epoch is not the current launch's reads as "not ready", so the state array is never cleared between launches
// (launch_stage2 zeroes it when the buffer is new and when the 22-bit epoch wraps); tickets are consumed relative to a
// per-launch base for the same reason.
constexpr unsigned long long kS2FlagAgg = 1ull << 62, kS2FlagInc = 2ull << 62, kS2ValMask = (1ull << 40) - 1;
constexpr uint32_t kS2EpochMask = (1u << 22) - 1;
__device__ __forceinline__ unsigned long long s2_pack(unsigned long long flag, uint32_t epoch, long long value) {
  return flag | ((unsigned long long)epoch << 40) | ((unsigned long long)value & kS2ValMask);
}
// flag of a state word as seen by launch `epoch` (0 = not ready / stale)
__device__ __forceinline__ uint32_t s2_flag(unsigned long long st, uint32_t epoch) {
  return (uint32_t(st >> 40) & kS2EpochMask) == epoch ? uint32_t(st >> 62) : 0u;
}

// Exclusive scan of the tile's 64 per-ray counts (-> s_off); publishes the tile aggregate right away so that the
// successors' look-backs can pass over this tile while it is still busy.  Returns the tile total.
__device__ __forceinline__ int s2_tile_scan(const int* s_cnt, int* s_off, int tile, int lane,
                                            unsigned long long* __restrict__ tile_state, uint32_t epoch) {
  const int a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
  int x = a + b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  const int excl = x - (a + b);
  s_off[2 * lane] = excl;
  s_off[2 * lane + 1] = excl + a;
  const int tile_total = __shfl_sync(0xffffffffu, x, 31);
  if (tile > 0 && lane == 0) atomicExch(&tile_state[tile], s2_pack(kS2FlagAgg, epoch, tile_total));
  return tile_total;
}

// Decoupled look-back over the preceding tiles (-> *s_prefix), then publishes this tile's inclusive prefix.
// The inclusive prefixes travel along the tile sequence one window per L2 round trip, and with ~10^4 small tiles
// that chain is the kernel's critical path: the window is 128 tiles (4 independent loads per lane in flight) so
// the chain is ~80 hops instead of ~320.
__device__ __forceinline__ void s2_lookback(long long* s_prefix, int tile, int tile_total, int n_tiles, int lane,
                                            unsigned long long* __restrict__ tile_state, long long* __restrict__ total,
                                            uint32_t epoch) {
  long long prefix = 0;
  if (tile > 0) {
    int idx = tile - 1;   // nearest predecessor not yet accounted for
    const long long t0 = clock64();
    bool done = false;
    while (!done) {
      unsigned long long st[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int my = idx - 32 * k - lane;
        st[k] = (my >= 0) ? ld_volatile_u64(&tile_state[my]) : s2_pack(kS2FlagInc, epoch, 0);   // virtual predecessor of tile 0
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (done) break;
        int first_inc;
        while (true) {
          const uint32_t fl = s2_flag(st[k], epoch);
          const uint32_t ready = __ballot_sync(0xffffffffu, fl != 0);
          const uint32_t inc = __ballot_sync(0xffffffffu, fl == 2);
          first_inc = inc ? (__ffs(inc) - 1) : 32;   // lanes [0, first_inc] must all be ready
          const uint32_t need = (first_inc >= 31) ? 0xffffffffu : ((2u << first_inc) - 1u);
          if ((ready & need) == need) break;
          if (clock64() - t0 > ADN_WATCHDOG_CYCLES) asm volatile("trap;");
          const int my = idx - 32 * k - lane;
          if (my >= 0) st[k] = ld_volatile_u64(&tile_state[my]);
        }
        long long contrib = (lane <= first_inc) ? (long long)(st[k] & kS2ValMask) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        prefix += contrib;
        if (first_inc < 32) done = true;
      }
      idx -= 128;
    }
  }
  if (lane == 0) {
    atomicExch(&tile_state[tile], s2_pack(kS2FlagInc, epoch, prefix + tile_total));
    *s_prefix = prefix;
    if (tile == n_tiles - 1) *total = prefix + tile_total;
  }
}

__global__ void __launch_bounds__(kS2Threads)
stage2_kernel(const float* __restrict__ raw0, long long n_rays, float thr, int K, const float* __restrict__ zlut,
              int32_t* __restrict__ count, int32_t* __restrict__ offset, int32_t* __restrict__ cell_out,
              int32_t* __restrict__ ray_out, float* __restrict__ z_out, float* __restrict__ zp_out,
              long long* __restrict__ total, unsigned long long* __restrict__ tile_state, unsigned int* __restrict__ ticket,
              int n_tiles, uint32_t epoch, uint32_t ticket_base) {
  __shared__ uint32_t s_sel[kS2Rays][4];
  __shared__ int s_cnt[kS2Rays];
  __shared__ int s_off[kS2Rays];
  __shared__ long long s_prefix;
  __shared__ int s_tile;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_tile = int(atomicAdd(ticket, 1u) - ticket_base);
  __syncthreads();
  const int tile = s_tile;
  const long long ray0 = (long long)tile * kS2Rays;

  // phase 1: selection (the warp's 8 row loads are issued up front: 8 x 512 B in flight per warp)
  float4 rows8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long r = ray0 + warp * 8 + i;
    rows8[i] = (r < n_rays) ? __ldg(reinterpret_cast<const float4*>(raw0 + r * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = warp * 8 + i;
    const long long r = ray0 + rl;
    int cnt = 0;
    uint32_t sel[4] = {0, 0, 0, 0};
    if (r < n_rays) cnt = select_cells(rows8[i], thr, K, lane, sel);
    if (lane == 0) {
      s_cnt[rl] = cnt;
#pragma unroll
      for (int j = 0; j < 4; ++j) s_sel[rl][j] = sel[j];
    }
  }
  __syncthreads();

  // CTA scan + decoupled look-back (warp 0)
  if (warp == 0) {
    const int tile_total = s2_tile_scan(s_cnt, s_off, tile, lane, tile_state, epoch);
    s2_lookback(&s_prefix, tile, tile_total, n_tiles, lane, tile_state, total, epoch);
  }
  __syncthreads();
  const long long prefix = s_prefix;

  // phase 2: write the packed samples (ray-major, ascending cell)
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rl = warp * 8 + i;
    const long long r = ray0 + rl;
    if (r >= n_rays) break;
    const float4 v4 = rows8[i];
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    const long long off = prefix + s_off[rl];
    if (lane == 0) {
      count[r] = s_cnt[rl];
      offset[r] = int32_t(off);
    }
    const uint32_t below = (1u << lane) - 1u;
    int rank = 0;
    uint32_t sel[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      sel[j] = s_sel[rl][j];
      rank += __popc(sel[j] & below);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if ((sel[j] >> lane) & 1u) {
        const int cell = 4 * lane + j;
        const long long o = off + rank;
        z_out[o] = zlut[cell];
        zp_out[o] = v[j];
        if (cell_out) cell_out[o] = cell;
        ray_out[o] = int32_t(r);
        ++rank;
      }
    }
  }
}

// ---- thread-per-ray variant (K <= 16): the default path.
// The warp-per-ray kernel above spends ~480 warp instructions per ray once most rays overflow K (cross-lane pop rounds),
// which makes it issue bound at ~15 % of HBM.  Here every THREAD owns a ray, so each instruction advances 32 rays:
//   * the tile's 64 rows land in shared memory through cp.async (each warp fetches its own 32 rows, 512 B per
//     instruction); rows are padded to 33 x 16 B so that "thread t reads chunk c of row t" is bank-conflict free,
//   * a ray is 16 groups of 8 cells; the thread keeps the 16 group maxima (shared memory, 5 x 16 B per thread) and
//     runs up to K rounds of: max group (first on ties) -> first cell of that group holding the max -> record it,
//     overwrite it with -inf, refresh the group's maximum.  Round 0 is always taken (the arg-max fallback of
//     :748-749), later rounds only while the popped value is >= thr, so one loop covers "none", "<= K" and "> K"
//     survivors with the reference's order (value descending, ties lower cell first) and no per-case branches,
//   * picks are emitted in ascending cell order (rank = popcount of the selection mask below the cell) into a
//     shared staging area that aliases the dead rows, then copied out coalesced after the tile's look-back scan.
constexpr int kS2tRowBytes = 528;
constexpr int kS2tGmBytes = 0;    // group maxima live in registers
constexpr size_t kS2tSmemBytes = size_t(kS2Rays) * (kS2tRowBytes + kS2tGmBytes) + 128 * sizeof(float);

template <int KMAX>
__global__ void __launch_bounds__(kS2Rays)
stage2_thread_kernel(const float* __restrict__ raw0, long long n_rays, float thr, int K, const float* __restrict__ zlut,
                     int32_t* __restrict__ count, int32_t* __restrict__ offset, int32_t* __restrict__ cell_out,
                     int32_t* __restrict__ ray_out, float* __restrict__ z_out, float* __restrict__ zp_out,
                     long long* __restrict__ total, unsigned long long* __restrict__ tile_state,
                     unsigned int* __restrict__ ticket, int n_tiles, uint32_t epoch, uint32_t ticket_base) {
  extern __shared__ __align__(16) uint8_t s2t_smem[];
  uint8_t* rows = s2t_smem;
  uint8_t* gms = rows + kS2Rays * kS2tRowBytes;
  float* zl = reinterpret_cast<float*>(gms + kS2Rays * kS2tGmBytes);
  __shared__ int s_cnt[kS2Rays];
  __shared__ int s_off[kS2Rays];
  __shared__ long long s_prefix;
  __shared__ int s_tile;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) s_tile = int(atomicAdd(ticket, 1u) - ticket_base);
  __syncthreads();
  const int tile = s_tile;
  const long long ray0 = (long long)tile * kS2Rays;
  const long long r = ray0 + tid;
  const bool valid = r < n_rays;
  uint8_t* row = rows + tid * kS2tRowBytes;
  // each warp fetches the 32 rows its own threads own: one 512-byte row per cp.async instruction (16 B per lane,
  // coalesced), nothing staged in registers, and only a __syncwarp between the copies and their consumers.
  // (A bulk copy per thread serialises into a 32-iteration uniform-register waterfall.)
  {
    const float* src = raw0 + (ray0 + warp * 32) * 128 + lane * 4;
    const uint32_t dst = smem_u32(rows + (warp * 32) * kS2tRowBytes + lane * 16);
    const long long left = n_rays - (ray0 + warp * 32);
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
      if (i < left)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * kS2tRowBytes), "l"(src + i * 128) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  zl[tid] = zlut[tid];
  zl[tid + 64] = zlut[tid + 64];
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();

  const float NEG = __int_as_float(0xff800000);
  const float4* row4 = reinterpret_cast<const float4*>(row);
  auto max8 = [](const float4 a, const float4 b) {
    return fmaxf(fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)), fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
  };
  float g[16];   // group maxima, registers
#pragma unroll
  for (int q = 0; q < 16; ++q) g[q] = max8(row4[2 * q], row4[2 * q + 1]);

  uint32_t mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;   // selection mask over the 128 cells
  float mv[KMAX];                                 // popped values, pick order
  uint32_t cp[KMAX / 4];                          // popped cells, 8 bits each
#pragma unroll
  for (int j = 0; j < KMAX / 4; ++j) cp[j] = 0;
  int cnt = 0;
  bool active = valid;
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    mv[j] = 0.0f;
    if (j >= K) break;                                        // warp uniform
    if (!__any_sync(0xffffffffu, active)) break;              // warp uniform
    const float m = fmaxf(fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7]))),
                          fmaxf(fmaxf(fmaxf(g[8], g[9]), fmaxf(g[10], g[11])), fmaxf(fmaxf(g[12], g[13]), fmaxf(g[14], g[15]))));
    int gi = 0;
#pragma unroll
    for (int i = 15; i >= 0; --i) gi = (g[i] == m) ? i : gi;   // first group holding the maximum
    float* grp = reinterpret_cast<float*>(row) + 8 * gi;
    const float4 a = reinterpret_cast<const float4*>(grp)[0], b = reinterpret_cast<const float4*>(grp)[1];
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    bool found = false;
    int idx = 0;
    float nm = NEG;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const bool e = (x[i] == m) && !found;    // first cell of the group holding the maximum
      found = found || e;
      idx = e ? i : idx;
      nm = fmaxf(nm, e ? NEG : x[i]);          // the group's maximum once that cell is gone
    }
    grp[idx] = NEG;
#pragma unroll
    for (int i = 0; i < 16; ++i) g[i] = (i == gi) ? nm : g[i];
    const bool sel = active && (j == 0 || m >= thr);
    active = sel;
    const int cell = 8 * gi + idx;
    const uint32_t bit = sel ? (1u << (cell & 31)) : 0u;
    const int w = cell >> 5;
    mk0 |= (w == 0) ? bit : 0u;
    mk1 |= (w == 1) ? bit : 0u;
    mk2 |= (w == 2) ? bit : 0u;
    mk3 |= (w == 3) ? bit : 0u;
    mv[j] = m;
    cp[j >> 2] |= uint32_t(cell) << (8 * (j & 3));
    cnt += sel ? 1 : 0;
  }

  s_cnt[tid] = cnt;
  __syncthreads();   // every thread is done with its row: the row area becomes the staging area
  int tile_total_w0 = 0;
  if (warp == 0) tile_total_w0 = s2_tile_scan(s_cnt, s_off, tile, lane, tile_state, epoch);
  __syncthreads();
  const int off = s_off[tid];

  float* st_z = reinterpret_cast<float*>(rows);
  float* st_zp = st_z + kS2Rays * KMAX;
  int32_t* st_ray = reinterpret_cast<int32_t*>(st_zp + kS2Rays * KMAX);
  int32_t* st_cell = st_ray + kS2Rays * KMAX;
  const int c0 = __popc(mk0), c1 = c0 + __popc(mk1), c2 = c1 + __popc(mk2);
#pragma unroll
  for (int j = 0; j < KMAX; ++j) {
    if (j < cnt) {
      const int cell = int((cp[j >> 2] >> (8 * (j & 3))) & 127u);
      const int w = cell >> 5;
      const uint32_t mk = (w == 0) ? mk0 : (w == 1) ? mk1 : (w == 2) ? mk2 : mk3;
      const int base = (w == 0) ? 0 : (w == 1) ? c0 : (w == 2) ? c1 : c2;
      const int p = off + base + __popc(mk & ((1u << (cell & 31)) - 1u));
      st_z[p] = zl[cell];
      st_zp[p] = mv[j];
      st_ray[p] = int32_t(r);
      st_cell[p] = cell;
    }
  }
  // the look-back (a chain of L2 round trips) starts only after this warp's own staging work is out of the way
  if (warp == 0) s2_lookback(&s_prefix, tile, tile_total_w0, n_tiles, lane, tile_state, total, epoch);
  __syncthreads();
  const long long prefix = s_prefix;
  if (valid) {
    count[r] = cnt;
    offset[r] = int32_t(prefix + off);
  }
  const int tile_total = s_off[kS2Rays - 1] + s_cnt[kS2Rays - 1];
  for (int i = tid; i < tile_total; i += kS2Rays) {
    const long long o = prefix + i;
    z_out[o] = st_z[i];
    zp_out[o] = st_zp[i];
    ray_out[o] = st_ray[i];
    if (cell_out) cell_out[o] = st_cell[i];
  }
}

size_t stage2_scratch_bytes(long long n_rays) {
  const long long n_tiles = (n_rays + kS2Rays - 1) / kS2Rays;
  return size_t(n_tiles + 2) * 8;
}

cudaError_t launch_stage2(const float* d_raw0, long long n_rays, float thr, int K, const float* d_zlut, int32_t* d_count,
                          int32_t* d_offset, int32_t* d_cell, int32_t* d_ray, float* d_z, float* d_zp, long long* d_total,
                          void* d_scratch, Stage2Sync* sync, cudaStream_t s) {
  if (n_rays <= 0) return cudaMemsetAsync(d_total, 0, sizeof(long long), s);
  const int n_tiles = int((n_rays + kS2Rays - 1) / kS2Rays);
  const size_t bytes = stage2_scratch_bytes(n_rays);
  // [ticket | states]: cleared only when the buffer is new / has grown or the epoch wraps (see the state-word comment)
  sync->epoch = (sync->epoch + 1) & kS2EpochMask;
  if (sync->scratch != d_scratch || sync->cleared_bytes < bytes || sync->epoch == 0) {
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, bytes, s);
    if (e != cudaSuccess) return e;
    sync->scratch = d_scratch;
    sync->cleared_bytes = bytes;
    sync->epoch = 1;
    sync->ticket_base = 0;
  }
  unsigned int* ticket = reinterpret_cast<unsigned int*>(d_scratch);
  unsigned long long* state = reinterpret_cast<unsigned long long*>(d_scratch) + 1;
  const uint32_t epoch = sync->epoch, base = sync->ticket_base;
  sync->ticket_base += uint32_t(n_tiles);   // the launch consumes exactly n_tiles tickets (unsigned wrap-around is fine)
  // thread-per-ray kernel whenever its assumptions hold (K <= 16, 16-byte aligned rows for the cp.async fetch)
  const bool aligned = (reinterpret_cast<uintptr_t>(d_raw0) & 15u) == 0;
  if (K <= 8 && aligned)
    stage2_thread_kernel<8><<<n_tiles, kS2Rays, kS2tSmemBytes, s>>>(d_raw0, n_rays, thr, K, d_zlut, d_count, d_offset, d_cell, d_ray,
                                                                   d_z, d_zp, d_total, state, ticket, n_tiles, epoch, base);
  else if (K <= 16 && aligned)
    stage2_thread_kernel<16><<<n_tiles, kS2Rays, kS2tSmemBytes, s>>>(d_raw0, n_rays, thr, K, d_zlut, d_count, d_offset, d_cell, d_ray,
                                                                    d_z, d_zp, d_total, state, ticket, n_tiles, epoch, base);
  else
    stage2_kernel<<<n_tiles, kS2Threads, 0, s>>>(d_raw0, n_rays, thr, K, d_zlut, d_count, d_offset, d_cell, d_ray, d_z, d_zp,
                                                 d_total, state, ticket, n_tiles, epoch, base);
  return cudaGetLastError();
}

__global__ void stage2_dense_kernel(long long n_rays, int K, int32_t* __restrict__ count, int32_t* __restrict__ offset,
                                    long long* __restrict__ total) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i == 0) *total = n_rays * K;
  if (i >= n_rays) return;
  if (count) count[i] = K;
  if (offset) offset[i] = int32_t(i * K);
}

cudaError_t launch_stage2_dense(long long n_rays, int K, int32_t* d_count, int32_t* d_offset, long long* d_total,
                                cudaStream_t s) {
  const long long n = n_rays > 0 ? n_rays : 1;
  stage2_dense_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(n_rays, K, d_count, d_offset, d_total);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- stage 3
// RayMarchFromPoses.batch (src/features.py:458-479): one thread per packed sample.
__global__ void __launch_bounds__(128)
stage3_kernel(const __grid_constant__ SceneDev sc, const float* __restrict__ ray_o, const float* __restrict__ ray_d,
              const int32_t* __restrict__ ray_idx, const float* __restrict__ z, const float* __restrict__ zlut_dense, int K,
              long long n_samples_host, const long long* __restrict__ n_samples_dev, float* __restrict__ x1,
              uint8_t* __restrict__ tiles1) {
  const long long n_samples = n_samples_dev ? *n_samples_dev : n_samples_host;
  const long long n_pad = ((n_samples + kTileM - 1) / kTileM) * kTileM;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_pad; i += (long long)gridDim.x * blockDim.x) {
    float f[128];
#pragma unroll
    for (int j = 0; j < 128; ++j) f[j] = 0.0f;
    if (i < n_samples) {
      long long r;
      float zw;
      if (ray_idx) {
        r = ray_idx[i];
        zw = z[i];
      } else {
        r = i / K;
        zw = zlut_dense[i - r * K];
      }
      float o[3], d[3], pos[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        o[a] = __ldg(ray_o + 3 * r + a);
        d[a] = __ldg(ray_d + 3 * r + a);
      }
      if (sc.ndc) {
        // ndc_rays(H, W, focal, near = 1) (src/nerf_raymarch_common.py:71-88) in the reference's operation order, then
        // pos = o' + d' z with the un-normalised NDC direction, no position normalisation, view encoding of d' / |d'|
        const float t = __fdiv_rn(-__fadd_rn(1.0f, o[2]), d[2]);
        float on[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) on[a] = __fadd_rn(o[a], __fmul_rn(t, d[a]));
        const float q0 = __fdiv_rn(on[0], on[2]), q1 = __fdiv_rn(on[1], on[2]);
        const float o0 = __fdiv_rn(__fmul_rn(sc.ndc_cw, on[0]), on[2]);
        const float o1 = __fdiv_rn(__fmul_rn(sc.ndc_ch, on[1]), on[2]);
        const float o2 = __fadd_rn(1.0f, __fdiv_rn(2.0f, on[2]));
        const float d0 = __fmul_rn(sc.ndc_cw, __fsub_rn(__fdiv_rn(d[0], d[2]), q0));
        const float d1 = __fmul_rn(sc.ndc_ch, __fsub_rn(__fdiv_rn(d[1], d[2]), q1));
        const float d2 = __fdiv_rn(-2.0f, on[2]);
        pos[0] = __fadd_rn(o0, __fmul_rn(d0, zw));                                                    // :458
        pos[1] = __fadd_rn(o1, __fmul_rn(d1, zw));
        pos[2] = __fadd_rn(o2, __fmul_rn(d2, zw));
        const float dn = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
        d[0] = __fdiv_rn(d0, dn);                                                                     // :431
        d[1] = __fdiv_rn(d1, dn);
        d[2] = __fdiv_rn(d2, dn);
      } else {
#pragma unroll
        for (int a = 0; a < 3; ++a) pos[a] = __fsub_rn(__fadd_rn(o[a], __fmul_rn(d[a], zw)), sc.c[a]);   // :458, loc = pos - c
        // normalization_inverse_sqrt_dist_centered (src/nerf_raymarch_common.py:226-230)
        const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(pos[0], pos[0]), __fmul_rn(pos[1], pos[1])), __fmul_rn(pos[2], pos[2])));
        const float den = __fmul_rn(sc.sqrt_max_depth, __fsqrt_rn(nrm));
#pragma unroll
        for (int a = 0; a < 3; ++a) pos[a] = __fdiv_rn(pos[a], den);
      }
      posenc3<kNFreqPos>(pos, f);                      // 63: position block FIRST (:473-479)
      posenc3<kNFreqDir>(d, f + 64);                   // 27 (un-renormalised nds), staged at column 64
      if (x1) {
#pragma unroll
        for (int j = 0; j < 63; ++j) x1[i * kFeat + j] = f[j];
#pragma unroll
        for (int j = 0; j < 27; ++j) x1[i * kFeat + 63 + j] = f[64 + j];
      }
    }
    if (tiles1) {
      // packed MLP1 input: per tile [P: 63 pos features + 0 | V: 27 dir features + zeros], 16 KB each; staged in
      // shared memory and written with one bulk shared -> global copy per tile (TMA engine)
      extern __shared__ __align__(1024) uint8_t s_tile[];
      const long long t = i >> 7;
      const uint32_t r = uint32_t(i & 127);
      if (threadIdx.x == 0) bulk_wait_read_all();   // the previous tile's copy has finished reading s_tile
      __syncthreads();
#pragma unroll
      for (int b = 0; b < 2; ++b) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const float* v = f + b * 64 + ch * 8;
          uint4 hi;
          hi.x = bf16x2(v[0], v[1]);
          hi.y = bf16x2(v[2], v[3]);
          hi.z = bf16x2(v[4], v[5]);
          hi.w = bf16x2(v[6], v[7]);
          *reinterpret_cast<uint4*>(s_tile + b * kBlkBytes + sw128_offset(r, uint32_t(ch * 8))) = hi;
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (threadIdx.x == 0) {
        bulk_s2g(tiles1 + size_t(t) * (2 * kBlkBytes), s_tile, 2 * kBlkBytes);
        bulk_commit();
      }
    }
  }
  if (tiles1 && threadIdx.x == 0) bulk_wait_all();
}

cudaError_t launch_stage3(const SceneDev& sc, const float* d_ray_o, const float* d_ray_d, const int32_t* d_ray,
                          const float* d_z, const float* d_zlut_dense, int K, long long n_samples, const long long* d_total,
                          float* d_x1, uint8_t* d_tiles1, cudaStream_t s) {
  // n_samples is an upper bound (capacity) when d_total is given
  if (n_samples <= 0) return cudaSuccess;
  long long blocks = (n_samples + kTileM - 1) / kTileM;
  const long long cap = 148ll * 64;
  if (d_total && blocks > cap) blocks = cap;   // grid-stride when the true count lives on the device
  static unsigned long long attr_done = 0;   // per device
  if (set_max_dyn_smem_once(reinterpret_cast<const void*>(stage3_kernel), 2 * kBlkBytes, &attr_done) != cudaSuccess) return cudaGetLastError();
  stage3_kernel<<<unsigned(blocks), 128, d_tiles1 ? size_t(2 * kBlkBytes) : 0, s>>>(sc, d_ray_o, d_ray_d, d_ray, d_z, d_zlut_dense, K,
                                                                                      n_samples, d_total, d_x1, d_tiles1);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- stage 5
// adaptive_raw2outputs (src/nerf_raymarch_common.py:91-144), accumulation_mult == "alpha".
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__device__ __forceinline__ uint32_t to_rgba8(float r, float g, float b) {
  // adaptive_cuda_kernels.cu:846-851: clamp to [0,1] * 255, alpha = 255
  const uint32_t R = uint32_t(fminf(fmaxf(r, 0.0f), 1.0f) * 255.0f);
  const uint32_t G = uint32_t(fminf(fmaxf(g, 0.0f), 1.0f) * 255.0f);
  const uint32_t B = uint32_t(fminf(fmaxf(b, 0.0f), 1.0f) * 255.0f);
  return R | (G << 8) | (B << 16) | (255u << 24);
}

// Per-ray epilogue shared by both composite kernels: acc / disparity / log-warped depth (src/nerf_raymarch_common.py:137-139,
// src/util/depth_transformations.py:15-35).
__device__ __forceinline__ void s5_write_ray_aux(const Stage5Aux& aux, long long r, float dm, float acc) {
  if (aux.depth_map) aux.depth_map[r] = dm;
  if (aux.acc_map) aux.acc_map[r] = acc;
  if (aux.disp_map) aux.disp_map[r] = __fdiv_rn(1.0f, fmaxf(1e-10f, __fdiv_rn(dm, acc)));
  if (aux.depth_est && aux.linear_depth) {
    aux.depth_est[r] = dm;
  } else if (aux.depth_est) {
    float d = __fsub_rn(dm, aux.dr_min);
    if (d <= 0.0f) d = 0.001f;
    aux.depth_est[r] = __fdiv_rn(logf(__fadd_rn(d, 1.0f)), aux.log_range);
  }
}

// One thread per ray, samples visited in order: the same sequential cumprod / sum order as torch.
__global__ void __launch_bounds__(128)
stage5_thread_kernel(const float4* __restrict__ raw1, const float* __restrict__ zp, const float* __restrict__ z,
                     const int32_t* __restrict__ offset, const int32_t* __restrict__ count, long long n_rays, int K,
                     float* __restrict__ rgb, uint32_t* __restrict__ rgba8, const Stage5Aux aux) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= n_rays) return;
  const long long off = offset[r];
  const int n = count[r];
  const bool want_z = aux.depth_map || aux.disp_map || aux.depth_est || aux.z_vals;
  float T = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, dm = 0.0f, acc = 0.0f;
  for (int j0 = 0; j0 < n; j0 += 4) {
    // four samples' loads are issued before the (sequential) transmittance chain consumes them
    float4 q4[4];
    float zp4[4], z4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool live = j0 + u < n;
      q4[u] = live ? __ldg(raw1 + off + j0 + u) : make_float4(0.f, 0.f, 0.f, 0.f);
      zp4[u] = live ? __ldg(zp + off + j0 + u) : 0.0f;
      z4[u] = (live && want_z) ? __ldg(z + off + j0 + u) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u;
      if (j >= n) break;
      const float4 q = q4[u];
      const float sr = sigmoidf_acc(q.x), sg = sigmoidf_acc(q.y), sb = sigmoidf_acc(q.z), sa = sigmoidf_acc(q.w);
      const float alpha = __fmul_rn(sa, zp4[u]);                                    // :123-125
      const float w = __fmul_rn(alpha, T);                                          // :128-129
      T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f));
      cr = __fadd_rn(cr, __fmul_rn(w, sr));                                         // :135
      cg = __fadd_rn(cg, __fmul_rn(w, sg));
      cb = __fadd_rn(cb, __fmul_rn(w, sb));
      acc = __fadd_rn(acc, w);                                                      // :139
      if (want_z) {
        dm = __fadd_rn(dm, __fmul_rn(w, z4[u]));                                    // :137
        if (aux.z_vals) aux.z_vals[r * K + j] = z4[u];
      }
      if (aux.weights) aux.weights[r * K + j] = w;
      if (aux.alpha) aux.alpha[r * K + j] = alpha;
    }
  }
  for (int j = n; j < K; ++j) {
    if (aux.weights) aux.weights[r * K + j] = 0.0f;
    if (aux.alpha) aux.alpha[r * K + j] = 0.0f;
    if (aux.z_vals) aux.z_vals[r * K + j] = __int_as_float(0x7fc00000);          // features.py:546 (NaN padding)
  }
  if (rgb) {
    rgb[3 * r + 0] = cr;
    rgb[3 * r + 1] = cg;
    rgb[3 * r + 2] = cb;
  }
  if (rgba8) rgba8[r] = to_rgba8(cr, cg, cb);
  s5_write_ray_aux(aux, r, dm, acc);
}

// One warp per ray (dense 128 samples / large K): lanes own consecutive samples, transmittance by a
// warp-wide product scan with a running carry.
__global__ void __launch_bounds__(256)
stage5_warp_kernel(const float4* __restrict__ raw1, const float* __restrict__ zp, const float* __restrict__ z,
                   const float* __restrict__ zlut_dense, const int32_t* __restrict__ offset,
                   const int32_t* __restrict__ count, long long n_rays, int K, int dense, float* __restrict__ rgb,
                   uint32_t* __restrict__ rgba8, const Stage5Aux aux) {
  const long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_rays) return;
  const long long off = dense ? r * K : (long long)offset[r];
  const int n = dense ? K : count[r];
  const bool want_z = aux.depth_map || aux.disp_map || aux.depth_est || aux.z_vals;
  float carry = 1.0f, cr = 0.0f, cg = 0.0f, cb = 0.0f, dm = 0.0f, acc = 0.0f;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const int j = j0 + lane;
    float alpha = 0.0f, sr = 0.0f, sg = 0.0f, sb = 0.0f, zz = 0.0f;
    if (j < n) {
      const float4 q = __ldg(raw1 + off + j);
      sr = sigmoidf_acc(q.x);
      sg = sigmoidf_acc(q.y);
      sb = sigmoidf_acc(q.z);
      alpha = __fmul_rn(sigmoidf_acc(q.w), __ldg(zp + off + j));
      if (want_z) zz = dense ? __ldg(zlut_dense + j) : __ldg(z + off + j);
    }
    float f = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    // inclusive product scan
    float p = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float y = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) p = __fmul_rn(p, y);
    }
    float excl = __shfl_up_sync(0xffffffffu, p, 1);
    if (lane == 0) excl = 1.0f;
    const float T = __fmul_rn(carry, excl);
    const float w = __fmul_rn(alpha, T);
    carry = __fmul_rn(carry, __shfl_sync(0xffffffffu, p, 31));
    cr = __fadd_rn(cr, __fmul_rn(w, sr));
    cg = __fadd_rn(cg, __fmul_rn(w, sg));
    cb = __fadd_rn(cb, __fmul_rn(w, sb));
    dm = __fadd_rn(dm, __fmul_rn(w, zz));
    acc = __fadd_rn(acc, w);
    if (j < K) {
      const bool live = j < n;
      if (aux.weights) aux.weights[r * K + j] = live ? w : 0.0f;
      if (aux.alpha) aux.alpha[r * K + j] = live ? alpha : 0.0f;
      if (aux.z_vals) aux.z_vals[r * K + j] = live ? zz : __int_as_float(0x7fc00000);
    }
  }
  for (int j = ((n + 31) & ~31) + lane; j < K; j += 32) {
    if (aux.weights) aux.weights[r * K + j] = 0.0f;
    if (aux.alpha) aux.alpha[r * K + j] = 0.0f;
    if (aux.z_vals) aux.z_vals[r * K + j] = __int_as_float(0x7fc00000);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o);
    cg += __shfl_xor_sync(0xffffffffu, cg, o);
    cb += __shfl_xor_sync(0xffffffffu, cb, o);
    dm += __shfl_xor_sync(0xffffffffu, dm, o);
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
  }
  if (lane == 0) {
    if (rgb) {
      rgb[3 * r + 0] = cr;
      rgb[3 * r + 1] = cg;
      rgb[3 * r + 2] = cb;
    }
    if (rgba8) rgba8[r] = to_rgba8(cr, cg, cb);
    s5_write_ray_aux(aux, r, dm, acc);
  }
}

cudaError_t launch_stage5(const float* d_raw1, const float* d_zp, const float* d_z, const float* d_zlut_dense,
                          const int32_t* d_offset, const int32_t* d_count, long long n_rays, int K, int dense, float* d_rgb,
                          uint8_t* d_rgba8, const Stage5Aux& aux, cudaStream_t s) {
  if (n_rays <= 0) return cudaSuccess;
  const float4* raw = reinterpret_cast<const float4*>(d_raw1);
  uint32_t* rgba = reinterpret_cast<uint32_t*>(d_rgba8);
  if (dense || K > 32) {
    const long long threads = n_rays * 32;
    stage5_warp_kernel<<<unsigned((threads + 255) / 256), 256, 0, s>>>(raw, d_zp, d_z, d_zlut_dense, d_offset, d_count, n_rays,
                                                                        K, dense, d_rgb, rgba, aux);
  } else {
    stage5_thread_kernel<<<unsigned((n_rays + 127) / 128), 128, 0, s>>>(raw, d_zp, d_z, d_offset, d_count, n_rays, K, d_rgb,
                                                                         rgba, aux);
  }
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- metrics
__global__ void __launch_bounds__(256)
sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, int clamp01,
                      double* __restrict__ partials) {
  __shared__ double s_w[8];
  double acc = 0.0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    float x = __ldg(a + i);
    if (clamp01) x = fminf(fmaxf(x, 0.0f), 1.0f);
    const double d = double(x) - double(__ldg(b + i));
    acc += d * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_w[w];
    partials[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32) sqdiff_final_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) acc += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) *out = acc;
}

cudaError_t launch_image_sqdiff(const float* d_a, const float* d_b, long long n_values, int clamp01, double* d_partials,
                                double* d_sum, cudaStream_t s) {
  sqdiff_partial_kernel<<<kMetricBlocks, 256, 0, s>>>(d_a, d_b, n_values, clamp01, d_partials);
  sqdiff_final_kernel<<<1, 32, 0, s>>>(d_partials, kMetricBlocks, d_sum);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------- viewer surface
__global__ void rgba_to_surface_kernel(const uchar4* __restrict__ px, int W, int row0, int rows, cudaSurfaceObject_t surf) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x < W && y < rows) surf2Dwrite(px[size_t(y) * W + x], surf, x * int(sizeof(uchar4)), row0 + y);
}

cudaError_t launch_rgba_to_surface(const uint8_t* d_rgba8, int W, int row0, int rows, unsigned long long surface, cudaStream_t s) {
  if (rows <= 0 || W <= 0) return cudaSuccess;
  const dim3 block(32, 8), grid(unsigned((W + 31) / 32), unsigned((rows + 7) / 8));
  rgba_to_surface_kernel<<<grid, block, 0, s>>>(reinterpret_cast<const uchar4*>(d_rgba8), W, row0, rows, cudaSurfaceObject_t(surface));
  return cudaGetLastError();
}

}  // namespace adn
