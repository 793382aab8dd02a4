// Schedule description shared by host (plan builder, api.cu) and device (mega.cu).
//
// The whole pose-lifting hot path (reference: model/bilinear.py:7-41 forward, autograd backward
// driven by train_bilinear.py:75-83, clip + Adam) is expressed as a list of PHASES executed by one
// persistent kernel; a phase is a set of GEMM tile jobs and/or elementwise jobs that have no
// dependencies among themselves, and consecutive phases are separated by a grid-wide barrier.
// Launching a single-phase schedule gives a per-layer kernel; launching the full list gives the
// persistent whole-network kernel (cooperative launch).
//
// Geometry convention for every GEMM job: D[m][n] = sum_k A[m][k] * B[n][k]
//   m  (TMEM lane / one epilogue thread per row) = an output FEATURE,
//   n  (TMEM column)                              = a BATCH index (fwd/dgrad) or an input feature (wgrad).
// A job may be split along K (`ksplit` slices, each slice its own tile):
//   * small batches (the whole batch is one n tile): every fwd / dgrad GEMM is cut into 16 K slices so
//     that ~128 SMs stream the weight matrix at once; the slices store fp32 partial sums
//     (EPI_PARTIAL) and the following elementwise phase adds them in a fixed order while it does the
//     BatchNorm / ReLU / dropout math with one WARP per feature (EW_BN_*_SMALL);
//   * large batches: the wgrad GEMMs (K = batch) are cut so that their tiles weigh about as much as
//     a few dgrad tiles; slices add into the gradient buffer with vector reductions.
// Tiles of a phase are assigned to CTAs by a host-computed longest-first schedule (Phase::sched_off).
#pragma once
#include <stdint.h>

namespace b200 {

constexpr int kHidden = 1024;
constexpr int kDin = 32;    // 2 * 16 joints (reference model/bilinear.py:20-22)
constexpr int kDout = 48;   // 3 * 16 joints (reference model/bilinear.py:29)
constexpr int kNumLinear = 6;
constexpr int kNumHeavy = 5;
constexpr float kBnEps = 1e-5f;  // nn.BatchNorm1d default

enum EpiKind : int32_t {
  EPI_STORE_F32 = 0,   // D -> out_f32[m][n]                       (debug / unit test)
  EPI_PARTIAL = 1,     // D -> part[ks][m][n] fp32 partial sums of K slice ks
  EPI_FWD_LARGE = 2,   // bias, write z, per-tile partial sum / sumsq -> stats
  EPI_FWD_EVAL = 3,    // scale*acc+shift, ReLU (+skip)            (BN folded, eval mode)
  EPI_DECODE = 4,      // bias -> p[b][48]; optional fused MSE loss + dP
  EPI_DECODE_T = 5,    // the same with the GEMM transposed: rows = poses, 48 columns (large-batch plans)
  EPI_DGRAD_LARGE = 6, // da (+skip grad) -> dy, partial sum(dy), sum(dy*xhat) -> stats
  EPI_WGRAD = 7,       // dW tile -> flat grad buffer + sum of squares (ksplit > 1: vector reductions)
};

enum EwKind : int32_t {
  EW_PREP = 0,         // zero accumulators, split x into feature-major planes, bump BN counters
  EW_RESERVED1 = 1,    // (dropout bits are drawn where they are consumed)
  EW_BN_FOLD = 2,      // eval: scale = g*rsqrt(rv+eps), shift = b + (bias-rm)*scale, all five layers (Args::fold)
  EW_BN_FWD_APPLY = 3, // large path: z -> h + keep bits + running stats
  EW_BN_BWD_APPLY = 4, // large path: dy,z -> dz + db
  EW_SPLIT_DP = 5,     // external dP [B][48] fp32 -> feature-major planes + db5 (autograd path)
  EW_SUMSQ = 6,        // sum of squares of the flat gradient
  EW_ADAM = 7,         // clip coefficient + Adam + refresh bf16 operand planes of the weights
  EW_SPLIT_W = 8,      // fp32 master weights -> bf16 operand planes, no update
  EW_BN_FWD_SMALL = 9, // small path: sum K-slice partials -> bias, BN (batch or running stats), ReLU, dropout, skip
  EW_BN_BWD_SMALL = 10,// small path: sum partials (+skip grad) -> dropout/ReLU/BN backward -> dz, dgamma, dbeta, db
  EW_DECODE_SMALL = 11,// small path: sum partials + bias -> p; optional MSE loss, dP planes, decode bias grad
  EW_COMM_BARRIER = 12,// data parallel: cross-GPU barrier through flags in peer memory (block 0 only)
  EW_ALLREDUCE = 13,   // data parallel: reduce-scatter + all-gather of the flat gradient over NVLink peer memory
  EW_DGRAD_STATS = 14, // split large bf16 plans: sums of dy, dy*xhat from the raw gradient a thin dgrad GEMM stored
};

// bf16 "plane" tensor: value = sum_p plane[p].  planes = 1 (bf16 mode) or 3 (fp32-parity mode:
// 3 x 8 significand bits carry a full fp32 value; the GEMM issues the 6 significant cross products).
struct PlaneRef {
  uint16_t* p;           // bf16 bits
  int64_t plane_stride;  // elements between planes
  int64_t ld;            // elements between rows
};

// Per-call values.  Schedules (jobs, tensor maps) are built once per plan and live in device memory;
// everything that changes from one call to the next travels in the kernel parameters instead.
struct Dyn {
  const float* x;          // [B][32] fp32 network input           (train_bilinear.py:62,72)
  const float* y;          // [B][48] fp32 regression target       (train_bilinear.py:63,73)
  float* p_out;            // [B][48] fp32 prediction (nullptr: not stored)
  const float* dP;         // [B][48] fp32 upstream gradient (autograd path)
  const uint8_t* masks;    // [5][B][1024] injected dropout keep masks (nullptr: Philox)
  float* loss_out;         // mean MSE of this step
  float* gnorm_out;        // pre-clip global gradient norm
  uint64_t seed, step;     // Philox key / counter salt
  float lr, beta1, beta2, adam_eps, bias_corr1, bias_corr2_sqrt;
  float max_norm;          // clip_grad_norm_ max_norm (<= 0: no clipping)
  float grad_scale;        // 1 / world_size under data parallel
  float momentum[kNumHeavy];  // >= 0: BN momentum; < 0: cumulative average (momentum=None)
  uint32_t comm_epoch;     // data parallel: flag value of the first cross-GPU barrier of this launch
};

enum DynSrc : int32_t { SRC_NONE = 0, SRC_X = 1, SRC_DP = 2 };

// Chunk gating between an apply pass and the GEMM tiles of the same phase (helper-warp build).
constexpr int kGateShift = 11;
constexpr int kGateCols = 1 << kGateShift;   // 2048 batch columns = 32 k-blocks = one wgrad K slice
enum GateKind : int32_t {
  GATE_NONE = 0,
  GATE_N = 1,   // the tile's n range is a batch range (forward / dgrad GEMMs: activations are the B operand)
  GATE_K = 2,   // the tile's k range is a batch range (wgrad: dz is the K-major A operand)
  GATE_M = 3,   // the tile's m range is a batch range (transposed decode: activations are the A operand)
};

constexpr int kMaxZero = 8;

// BatchNorm folding table for eval mode (one per plan, in device memory; EW_BN_FOLD via Args::fold).
struct alignas(16) FoldTab {
  const float* bias[kNumHeavy];
  const float* gamma[kNumHeavy];
  const float* beta[kNumHeavy];
  const float* run_mean[kNumHeavy];
  const float* run_var[kNumHeavy];
  float* scale[kNumHeavy];
  float* shift[kNumHeavy];
};

// Data-parallel exchange over NVLink / NVSwitch peer memory (one per plan, in device memory; the
// EW_COMM_BARRIER / EW_ALLREDUCE jobs and the Adam phase reach it through Args::comm).  Every rank maps
// every other rank's flat gradient buffer and control block (CUDA IPC); rank r owns slice r of the
// buffer: it sums that slice over all ranks and writes the sum back into every rank's buffer.
constexpr int kMaxRanks = 8;
struct CommCtl {                 // lives at the end of each rank's communication buffer
  uint32_t flag[kMaxRanks];      // flag[r] is written by rank r (release.sys) and polled by the owner
  uint32_t pad[8];
  double sq[kMaxRanks];          // sq[r]: sum of squares of the reduced slice r, written by rank r
};
struct alignas(16) CommTab {
  float* grads[kMaxRanks];       // peer-mapped flat gradient buffers, index = rank
  CommCtl* ctl[kMaxRanks];
  int32_t rank, world;
  int64_t numel;
  double* sq_local;              // this rank's accumulator for its slice (cleared at the start of a step)
};

// Adam / weight-plane table (one per plan, in device memory; referenced by EW_ADAM / EW_SPLIT_W /
// EW_SUMSQ jobs through Args::adam).
struct alignas(16) AdamTab {
  float* params;           // flat fp32 master parameters (reference parameters() order)
  float* grads;            // flat fp32 gradients
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  int64_t w_off[kNumLinear];   // offsets of the six Linear weights in the flat buffer
  int32_t w_rows[kNumLinear];
  int32_t w_cols[kNumLinear];
  uint16_t* planes;        // bf16 operand planes of the WHOLE flat buffer: plane q of parameter e lives at
  int64_t plane_stride;    // planes[q * plane_stride + e]; the weight matrices are TMA-mapped in place
};

struct alignas(16) Args {
  // ---- generic
  int32_t layer;           // 0..5
  int32_t planes;          // 1 | 3
  int32_t B;               // batch rows in this launch
  int32_t fuse_loss;       // decode: 1 = fused MSE loss + dP
  // ---- fp32 outputs / inputs
  float* out_f32;          // STORE: D; PARTIAL: slice 0 of the partial sums; WGRAD: dW [out][in]
  int64_t ld_out;
  int64_t part_stride;     // PARTIAL / EW_*_SMALL: elements between K slices
  const float* part;       // EW_*_SMALL: partial sums [nsplit][rows][ld_out]
  int32_t nsplit;          // EW_*_SMALL: number of K slices to add
  int32_t eval_mode;       // EW_BN_FWD_SMALL: 1 = running statistics (folded scale/shift), no dropout
  const float* bias;       // Linear bias [M]
  const float* gamma;      // BN weight  [M]
  const float* beta;       // BN bias    [M]
  float* run_mean;         // BN running_mean [M]   (updated in train)
  float* run_var;          // BN running_var  [M]
  const int64_t* nbt;      // BN num_batches_tracked (already incremented by EW_PREP)
  float* save_mean;        // [M] batch mean saved for backward
  float* save_rstd;        // [M] batch 1/sqrt(var+eps)
  double* stats;           // [M][2] accumulators (large path fwd: sum z, sum z^2; bwd: sum dy, sum dy*xhat)
  const float* scale;      // eval: folded BN scale [M]
  const float* shift;      // eval: folded BN shift [M]
  // ---- feature-major scratch ([f][ldT])
  void* zT;                // pre-BN activations, fp32 (planes==3) or bf16 (planes==1)
  void* dyT;               // large path: masked upstream grad, same dtype rule as zT
  int64_t ldT;             // row stride (elements) of zT / dyT / G*
  uint32_t* keep;          // [f][ldw] bit j of word w <-> b = 32 w + j : dropout_keep & (y > 0)
  const uint32_t* drop;    // [f][ldw] dropout keep bits (nullptr => all ones)
  int64_t ldw;
  const void* skip_in;     // backward: gradient arriving over the residual edge [f][ldT] (nullptr: none); fp32 on
  void* g_out;             // the small path, dtype of zT on the large path.  g_out: total gradient at this
                           // layer's output, kept for a later skip (large path: replaces dyT for that layer)
  PlaneRef skip_fm;        // fwd: residual input, feature-major planes (p == nullptr: none)
  // ---- plane output: feature-major [f][b] (every activation / gradient operand lives in this one
  // layout: K-major A operand of the wgrad GEMMs, MN-major B operand of the forward / dgrad GEMMs)
  PlaneRef out_fm;
  // ---- decode / loss
  double* loss_acc;        // sum (p-y)^2
  float* grad_bias;        // db   (decode: [48]; heavy: [M])
  float* grad_gamma;       // dgamma [M]
  float* grad_beta;        // dbeta  [M]
  double* gnorm2;          // sum g^2 accumulator
  // ---- elementwise extras
  int32_t src;             // DynSrc: which per-call tensor EW_PREP / EW_SPLIT_DP splits into planes
  int32_t src_cols;        // EW_PREP: columns of the source; EW_SUMSQ: 1 = non-weight tensors only
  void* zero_ptr[kMaxZero];     // EW_PREP: regions to clear (16-byte multiples)
  int64_t zero_bytes[kMaxZero];
  int64_t* nbt_w[kNumHeavy];    // EW_PREP / EW_DECODE_SMALL: BN counters to increment (nullptr: skip)
  const float* w_src;      // EW_BN_FWD_SMALL without partials: fp32 master weight [1024][32] of the first Linear
  float* x_save;           // small path: copy of x [B][32] kept by the forward for the fused first-layer wgrad
  float* grad_w0;          // EW_BN_BWD_SMALL of the first layer: dW0 [1024][32] (nullptr: no fused wgrad)
  const AdamTab* adam;     // EW_ADAM / EW_SPLIT_W / EW_SUMSQ
  const FoldTab* fold;     // EW_BN_FOLD
  const CommTab* comm;     // EW_COMM_BARRIER / EW_ALLREDUCE; EW_ADAM: clip norm from the exchanged slice sums
  int32_t comm_step;       // EW_COMM_BARRIER: 0 = before the exchange, 1 = after it (publishes the slice sum)
  int32_t pad_c;
  // ---- overlap of the large-batch BatchNorm apply passes with the GEMM that consumes their output (helper-warp
  // build, mega_pairx.cu): the apply pass walks the batch in chunks of kGateCols columns, all 1024 features of a
  // chunk before the next chunk, and counts finished (feature, chunk) items in gate[chunk]; GEMM tiles whose
  // operand lies in that chunk start once the counter reaches kHidden (GemmJob::gate)
  uint32_t* gate;          // EW_BN_FWD_APPLY / EW_BN_BWD_APPLY: per-chunk arrival counters (nullptr: no consumer waits)
  // ---- bf16 large-batch GEMM epilogues: tensor map (device address; box = 32 entries x 32 feature rows, 64-byte
  // swizzle) of the feature-major tensor the epilogue writes -- z (EPI_FWD_LARGE), g / dy (EPI_DGRAD_LARGE).  An
  // epilogue thread owns one feature row: its 64-byte runs go through a shared-memory staging tile and out with
  // ONE TMA store per warp and chunk instead of 32 two-sector requests through the load/store unit.  nullptr: direct stores.
  const void* tm_out;
};
constexpr int kEpiStageBytes = 16 * 1024;   // 4 epilogue warps x 2 staging tiles x 2 KB, at the top of the ring region

struct alignas(16) GemmJob {
  int32_t tmap_a, tmap_b;  // indices into the tensor-map table
  int32_t a_mn, b_mn;      // 1: operand is MN-major in shared memory (global layout [k][mn])
  int32_t M, N, K;
  int32_t BN;              // tile width == UMMA N (multiple of 16, <= 256)
  int32_t m_tiles, n_tiles, num_tiles;   // num_tiles = m_tiles * n_tiles * ksplit
  int32_t ksplit;          // >= 1: number of K slices (tile index = (ks * n_tiles + nt) * m_tiles + mt)
  int32_t kb_per_split;    // 64-element k-blocks per slice
  int32_t epi;
  int32_t gate_kind;       // GateKind: which extent of the tile lies along the batch (the gated dimension)
  int32_t pad_[2];
  const uint32_t* gate;    // per-chunk counters written by the apply pass running in the same phase (nullptr: none)
  Args a;
};

struct alignas(16) EwJob {
  int32_t kind;
  int32_t pad[3];
  Args a;
};

struct Phase {
  int32_t num_gemm, gemm_tiles;   // GEMM jobs of this phase (PhaseBlob::gj[0..num_gemm)) and their tile total
  int32_t num_ew;                 // elementwise jobs (PhaseBlob::ej[0..num_ew)), run after the CTA's tiles
  int32_t stage_bytes;            // smem ring geometry of this phase (max over its GEMM jobs)
  int32_t num_stages;
  int32_t sched_off;              // < 0: tiles are strided over CTAs; else offset into MegaParams::sched of the
                                  // tile-id lists (CtaTiles::begin/end index into it)
  int32_t pf_slice, pf_slices;    // pf_slices > 0: during this phase an idle warp prefetches slice pf_slice of
                                  // pf_slices of the optimizer's p / m / v arrays into L2 (MegaParams::adam)
  int32_t ew_first;               // helper-warp build: the elementwise jobs run CONCURRENTLY with the GEMM tiles of
                                  // the phase (on their own warps); elsewhere they run after the CTA's tiles
  int32_t aux_off;                // helper-warp build: byte offset inside the ring region of >= kAuxBytes that the
                                  // operand ring of this phase leaves free (per-feature tables of the apply pass); < 0: none
};
constexpr int kAuxBytes = 16 * 1024;

// Everything a CTA needs to know about one phase, contiguous in device memory.  All 256 threads copy
// the NEXT phase's blob into shared memory while thread 0 waits in the grid barrier, so no role ever
// chases descriptor pointers through L2 at the start of a phase.
constexpr int kBlobGemm = 2;
constexpr int kBlobEw = 3;
struct alignas(16) PhaseBlob {
  Phase ph;
  GemmJob gj[kBlobGemm];
  EwJob ej[kBlobEw];
};

// Tiles of one phase owned by one CTA (host-computed, longest first): ids [begin, end) of the phase's
// list in MegaParams::sched; the first kTileSlots ids are repeated inline.
constexpr int kTileSlots = 6;
struct alignas(16) CtaTiles {
  int32_t begin, end;
  int32_t tile[kTileSlots];
};

struct MegaParams {
  const PhaseBlob* blobs;   // [num_phases]
  const CtaTiles* cta_tiles;  // [num_phases][gridDim.x] (nullptr: every phase is strided)
  const void* tmaps;        // array of CUtensorMap (128 B each, 64 B aligned)
  const int32_t* sched;     // per-phase tile lists (see Phase::sched_off)
  const AdamTab* adam;      // optimizer arrays for the L2 prefetch (Phase::pf_slices), or nullptr
  unsigned int* barrier;    // grid barrier counter: monotonic across launches (never reset), see bar_base
  unsigned long long* phase_ns;  // [8 * (num_phases + 1)] globaltimer stamps written by block 0 (nullptr: off)
  int32_t num_phases;
  uint32_t dbg;             // timing experiments only (B200_DBG): 1 = skip the GEMM tiles, 2 = skip the apply passes and gates
  uint32_t bar_base;        // value of *barrier when this launch starts (the host keeps count)
  Dyn dyn;
};

}  // namespace b200
