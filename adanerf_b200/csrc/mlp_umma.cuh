// Fused multi-layer perceptron on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One persistent CTA per SM walks 128-row tiles of the batch through ALL layers of the network:
//   * weights stream layer by layer from L2 into a shared-memory ring with 1-D bulk async copies
//     (TMA engine, SASS UBLKCP), pre-packed on the host as K-major SWIZZLE_128B tiles,
//   * activations never leave the SM: the fp32 accumulator lives in TMEM, the epilogue warps read it
//     with tcgen05.ld, add bias / ReLU, round to bf16 (optionally a hi+lo split) and write the next
//     layer's A operand straight into swizzled shared memory,
//   * warp roles: warp 0 = weight producer, warp 1 = MMA issuer (one elected thread), warps 2..17 =
//     epilogue.  With NG = 2 two tiles ping-pong per CTA (tile A's epilogue overlaps tile B's MMAs).
//
// NSPLIT = 2 is the split-precision mode of the sampling network: x = hi + lo (both bf16) for
// activations and weights and three MMAs per K step (hi*hi + lo*hi + hi*lo), fp32 accumulate --
// fp32-class accuracy, which the bit-exact threshold decisions downstream need.
#pragma once
#include <cstdint>

namespace adn {

constexpr int kTileM = 128;
constexpr int kBlkBytes = 16384;  // one [128 x 64] bf16 SWIZZLE_128B block
constexpr int kMaxLayers = 12;
constexpr int kMlpThreads = 608;  // 19 warps: 16 epilogue, weight producer, barrier helper, MMA issuer
constexpr int kSideFloats = 3208; // fp32 side parameters (biases, alpha / rgb heads) carried in the kernel parameters

// Fixed side-parameter layout of the shading net for mlp_sh_kernel (mlp_sh.cu): every epilogue is specialised per layer,
// so its bias / head constants are compile-time offsets into MlpProgram::side and reach the FADD / FFMA as immediate
// constant-bank operands (no load instruction at all).  Layer l's fp32 bias: [256 l, 256 l + n_out).
constexpr int kShLayers = 10;
constexpr int kShAlphaW = 2560;   // alpha_linear.weight [256]
constexpr int kShRgbW = 2816;     // rgb_linear.weight [3][128]
constexpr int kShAlphaB = 3200;   // alpha_linear.bias
constexpr int kShRgbB = 3201;     // rgb_linear.bias [3]

enum : uint8_t {
  LF_RELU = 1,
  LF_ALPHA_DOT = 2,      // accumulate alpha = <post-activation row, alpha_w> on CUDA cores (fp32)
  LF_OUT_ACT = 4,        // write bf16 activations for the next layer
  LF_FINAL_RAW = 8,      // write fp32 rows to global (sampling net output / test programs)
  LF_FINAL_RGB = 16,     // rgb_linear on CUDA cores + write float4 (rgb, alpha)
  LF_LOAD_IN1_AFTER = 32,  // once this layer's MMAs are done, fetch the 2nd input block (view dirs)
  LF_WAIT_IN = 64          // the MMA warp must wait for that 2nd input before this layer
};

struct MlpLayer {
  uint32_t w_off;     // byte offset of this layer's packed weight stages (consumption order)
  uint32_t bias_off;  // float offset of the bias vector in MlpProgram::side
  uint8_t n_kb;       // number of 64-wide K blocks
  uint8_t a_blk[6];   // activation block index per K block
  uint8_t n_half;     // N / 128  (1 or 2)
  uint8_t flags;
  uint8_t out_blk0;   // first activation block the epilogue writes
  // mlp_sh_kernel: the (N half, stage) of this layer after whose MMAs the blocks that half 0's epilogue overwrites in
  // place (out_blk0, out_blk0 + 1) are no longer read -- the issuer commits `lo_free` there.  0xFF: not used.
  uint8_t lo_h, lo_s;
  // mlp_hp_kernel: 16-wide K steps issued per K block (4 = the whole 64-wide block; fewer when the tail columns of the
  // block are zero padding, e.g. 90 input features -> blocks of 4 and 2 steps; 30 features -> 2 and 0).
  uint8_t k_cnt[6];
};

struct MlpProgram {
  int32_t n_layers;
  int32_t in0_blk, in0_nblk;  // tile-start input: destination block, number of blocks (per term)
  int32_t in1_blk;            // 2nd input destination block
  int32_t hid_blk0;           // first hidden-activation block (blocks below it hold tile inputs)
  uint32_t in_tile_stride;    // bytes per tile in the packed input buffer
  uint32_t in0_off, in0_lo_off, in1_off;
  uint32_t alpha_w_off, alpha_b_off, rgb_w_off, rgb_b_off;  // float offsets in `side`
  int32_t out_cols;           // row stride of the FINAL_RAW output
  // The packed weight blob is replicated w_copies times in global memory, w_stride bytes apart; CTA pair p streams copy
  // p % w_copies.  Every CTA of a persistent launch re-reads the same ~1 MB every tile, nearly in lock step: with one copy
  // the whole grid hammers the same few dozen L2 slices at a time and the fills queue up behind each other.
  uint32_t w_copies, w_stride;
  MlpLayer layers[kMaxLayers];
  // mlp_sh_kernel: the issue schedule, one word per (layer, N half, stage) step in issue order, precomputed on the host so
  // that the MMA issuer and the dependency helper decode instead of deriving it (the issuer's instruction count per
  // 8-MMA group is the kernel's clock).  Bits: 0-3 / 4-7 activation block of K block 2s / 2s+1; 8 stage holds two K
  // blocks; 9-10 dependencies a slot picks up here for the first time in this layer (9: previous layer's half-0 epilogue,
  // 10: half-1 epilogue); 12 commit acc_full after this step; 13 commit lo_free; 14 N half; 15 first step of the half
  // (accumulator is overwritten, not accumulated); 16 the step's K block is the tile input (positions / view block), an
  // A operand that travels through the weight ring; 17 it is fetched at this step (two ring stages, one per slot, after
  // the weight stage); 18 released after this step; 19 it is the view block (two K steps); 20-21 stage index in the half.
  uint32_t sh_sched[kMaxLayers][6];
  uint8_t sh_steps[kMaxLayers];
  // Biases and the two tiny heads live in the kernel parameter (constant) bank: every lane of a warp reads
  // the same column's value, so the epilogue gets them through uniform constant loads, not the LSU.
  float side[kSideFloats];
};

// Describes how fp32 feature rows map onto the packed bf16 input blocks of a tile.
struct InputLayout {
  int32_t n_blk;
  int32_t src_col0[4];
  int32_t valid[4];
  uint32_t dst_off_hi[4];
  uint32_t dst_off_lo[4];
  uint32_t tile_stride;
  int32_t nsplit;
};


// Fused input encoder of the shading MLP (stage 3 inside the kernel): one extra warp computes the positional
// encoding of every packed sample (RayMarchFromPoses.batch, src/features.py:458-479) straight into the tile slot's
// input block, so the [M, 90] feature tensor never exists in HBM.  ray_idx == nullptr: dense mode (ray = sample / K,
// z = zlut_dense[sample % K]).
struct EncodeParams {
  const float* ray_o = nullptr;       // [N,3]
  const float* ray_d = nullptr;       // [N,3] (un-normalised, as SpherePosDir hands it on)
  const int32_t* ray_idx = nullptr;   // [M] packed sample -> ray
  const float* z = nullptr;           // [M] world depth of the packed samples
  const float* zlut_dense = nullptr;  // [K]
  int K = 1;
  float c[3] = {0.f, 0.f, 0.f};       // view_cell_center
  float sqrt_max_depth = 1.0f;
  uint8_t* scratch = nullptr;         // grid x 2 slots x 2 buffers x [P | V] blocks of 16 KB (mlp_enc_scratch_bytes)
};
size_t mlp_enc_scratch_bytes(int num_sms);
constexpr int kMlpEncThreads = kMlpThreads + 32;   // + 1 encoder warp (20 warps = 5 per SM sub-partition: still 96 registers)

// Shading net on mlp_sh_kernel (mlp_sh.cu): weight-stationary across two tile slots, N-half pipelined, CTA pairs.
// The program must be the one build_net1 emits in "sh" mode (fixed side layout, weights packed N half outermost).
cudaError_t launch_mlp_sh(const MlpProgram& prog, const uint8_t* wblob, const uint8_t* in_tiles, float* out,
                          const long long* rows_dev, long long rows_host, int* err_flag, int num_sms, cudaStream_t stream,
                          long long* trace = nullptr);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: set it once per (kernel, device).
// `done` is the caller's per-kernel bit mask (one static per launcher).
cudaError_t set_max_dyn_smem_once(const void* func, int bytes, unsigned long long* done);

// Launchers (defined in mlp_umma.cu).  rows_dev may be null (then rows_host is used).
cudaError_t launch_mlp(int nsplit, int ng, int cg, const MlpProgram& prog, const uint8_t* wblob,
                       const uint8_t* in_tiles, float* out, const long long* rows_dev, long long rows_host,
                       int* err_flag, int num_sms, cudaStream_t stream, long long* trace = nullptr,
                       const EncodeParams* enc = nullptr);
cudaError_t launch_pack_rows(const float* x, long long rows, const long long* rows_dev, int n_feat,
                             const InputLayout& lay, uint8_t* tiles, cudaStream_t stream);

}  // namespace adn
