// C ABI (include/bilinear_b200.h) + plan builder: carves the caller's workspace, encodes the TMA
// tensor maps and lays out the phase schedules that mega.cu executes.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>
#include <queue>
#include <atomic>
#include <cmath>
#include <string>
#include <vector>

#include "../../include/bilinear_b200.h"
#include "jobs.h"

namespace b200 {
cudaError_t launch_mega(const MegaParams& p, int grid, bool cooperative, cudaStream_t stream, int variant);
int mega_pair_max_clusters(int grid);
int mega_pairx_max_clusters(int grid);
cudaError_t launch_apply_fwd(const Args& a, const Dyn& d, cudaStream_t stream);
cudaError_t launch_apply_bwd(const Args& a, cudaStream_t stream);
cudaError_t launch_dgrad_stats(const Args& a, cudaStream_t stream);
int mega_ring_bytes();
struct EvalParams {
  const float* pred; const float* gt; const float* mean; const float* stddev; const int32_t* action;
  float* dist_out; float* pdist_out; double* acc_sum; double* pacc_sum; long long* acc_cnt;
  int64_t n; int32_t n_actions;
};
cudaError_t launch_eval(const EvalParams& prm, int joints, bool do_m, bool do_p, cudaStream_t stream);
cudaError_t launch_eval_tma(const EvalParams& prm, const CUtensorMap* tp, const CUtensorMap* tg, bool do_m, bool do_p,
                            cudaStream_t stream, int* launches);
int eval_tma_box_rows();
}  // namespace b200

using namespace b200;

// ------------------------------------------------------------------------------------------------
#ifdef B200_EPI_TMA
#define B200_EPI_TMA_HOST B200_EPI_TMA
#else
#define B200_EPI_TMA_HOST 0   // staged TMA-store epilogues are compiled out by default (mega.cu, B200_EPI_TMA)
#endif
static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}
#define CK(expr)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (expr);                                                                  \
    if (e_ != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// ------------------------------------------------------------------------------------------------
// TMA tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 plane tensor viewed as {inner, rows, planes}; box {64 inner elements (128 B), box_rows, 1};
// 128-byte swizzle; out-of-bounds elements read as zero (ragged batch / K tails / 48-row decode).
static int encode_map(CUtensorMap* out, const void* base, int64_t inner, int64_t rows, int planes, int64_t ld,
                      int64_t plane_stride, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  if (planes == 1) plane_stride = std::max<int64_t>(plane_stride, rows * ld);
  cuuint64_t dims[3] = {cuuint64_t(inner), cuuint64_t(rows), cuuint64_t(planes)};
  cuuint64_t strides[2] = {cuuint64_t(ld) * 2, cuuint64_t(plane_stride) * 2};
  cuuint32_t box[3] = {64, cuuint32_t(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((strides[0] % 16) || (strides[1] % 16) || (reinterpret_cast<uintptr_t>(base) % 16))
    return fail("tensor map alignment: ld=%lld plane_stride=%lld base=%p", (long long)ld, (long long)plane_stride, base);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled -> %d (inner=%lld rows=%lld planes=%d ld=%lld box_rows=%d)", int(r),
                (long long)inner, (long long)rows, planes, (long long)ld, box_rows);
  return 0;
}

// Output map of a bf16 large-batch epilogue (Args::tm_out): feature-major [rows][inner] tensor, box = 32 entries
// (64 bytes) x 32 rows, 64-byte swizzle -- the staging tile stage_store32_bf16 (mega.cu) fills.
static int encode_out_map(CUtensorMap* out, const void* base, int64_t inner, int64_t rows, int64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cuuint64_t(inner), cuuint64_t(rows)};
  cuuint64_t strides[1] = {cuuint64_t(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  if ((strides[0] % 16) || (reinterpret_cast<uintptr_t>(base) % 16))
    return fail("output tensor map alignment: ld=%lld base=%p", (long long)ld, base);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled (output) -> %d (inner=%lld rows=%lld ld=%lld)", int(r), (long long)inner,
                (long long)rows, (long long)ld);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// flat parameter layout (order of BilinearUnit.parameters(), model/bilinear.py:22-29)
// ------------------------------------------------------------------------------------------------
struct ParamLayout {
  int64_t w[kNumLinear], b[kNumLinear], g[kNumHeavy], be[kNumHeavy];
  int rows[kNumLinear], cols[kNumLinear];
  int64_t numel;
  ParamLayout() {
    int64_t o = 0;
    for (int l = 0; l < kNumLinear; ++l) {
      rows[l] = (l == 5) ? kDout : kHidden;
      cols[l] = (l == 0) ? kDin : kHidden;
      w[l] = o; o += int64_t(rows[l]) * cols[l];
      b[l] = o; o += rows[l];
      if (l < kNumHeavy) { g[l] = o; o += kHidden; be[l] = o; o += kHidden; }
    }
    numel = o;
  }
};
static const ParamLayout kPL;

// ------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------
struct Carver {
  uint8_t* base;
  int64_t off = 0;
  explicit Carver(void* b) : base(reinterpret_cast<uint8_t*>(b)) {}
  void* take(int64_t bytes) {
    off = (off + 1023) & ~int64_t(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  template <class T> T* take_n(int64_t n) { return reinterpret_cast<T*>(take(n * int64_t(sizeof(T)))); }
};

constexpr int kMaxTmaps = 192, kMaxPhases = 160, kMaxCtas = 192;
constexpr int kMaxSched = 1 << 18;   // int32 entries of per-phase tile schedules
constexpr int kSmallSplit = 16;      // K slices of the small-batch GEMMs (K = 1024 -> one k-block each)

struct Bufs {
  CUtensorMap* tmaps; PhaseBlob* blobs; CtaTiles* cta_tiles; unsigned int* barrier;
  unsigned long long* phase_ns;
  int32_t* sched; AdamTab* adam_tab; FoldTab* fold_tab; CommTab* comm_tab;
  uint16_t* wplanes; int64_t wplane_stride;   // bf16 planes of the whole flat parameter buffer
  float* part;                       // small path: [kSmallSplit][1024][Bs] fp32 partial sums
  float* x_save;                     // small path: copy of the input batch [Bs][32] fp32
  float* x_stage; float* y_stage; float* p_stage; float* scalars;   // device ends of the *_host entry points
  uint8_t* acc_base; int64_t acc_bytes;
  uint32_t* gates_f[kNumHeavy]; uint32_t* gates_b[kNumHeavy];   // helper-warp build: per-chunk counters of the apply passes
  int64_t gate_words;
  double* loss_acc; double* gnorm2; double* stats_f[kNumHeavy]; double* stats_b[kNumHeavy];
  float* save_mean[kNumHeavy]; float* save_rstd[kNumHeavy]; float* scale[kNumHeavy]; float* shift[kNumHeavy];
  PlaneRef W[kNumLinear];
  PlaneRef X_fm;
  void* zT[kNumHeavy]; void* dyT[kNumHeavy];
  uint32_t* keep[kNumHeavy]; uint32_t* drop;
  PlaneRef h_fm[kNumHeavy], dz_fm[kNumHeavy];
  void* G2; void* G4;
  PlaneRef dP_fm;
  int64_t Bs, ldw;
};

static PlaneRef take_planes(Carver& c, int planes, int64_t rows, int64_t ld) {
  PlaneRef r;
  const int64_t ps = (rows * ld + 511) & ~int64_t(511);
  r.p = c.take_n<uint16_t>(ps * planes);
  r.plane_stride = ps;
  r.ld = ld;
  return r;
}

static int64_t carve(Carver& c, Bufs& u, int B, int P, int kind) {
  const bool train = (kind == B200_PLAN_TRAIN);
  const bool large = B > (P == 1 ? 256 : 128);
  const int64_t Bs = (int64_t(B) + 31) & ~int64_t(31);
  const int64_t ldw = Bs / 32;
  u.Bs = Bs; u.ldw = ldw;
  u.tmaps = c.take_n<CUtensorMap>(kMaxTmaps);
  u.blobs = c.take_n<PhaseBlob>(kMaxPhases);
  u.cta_tiles = c.take_n<CtaTiles>(int64_t(kMaxPhases) * kMaxCtas);
  u.barrier = c.take_n<unsigned int>(64);
  u.phase_ns = c.take_n<unsigned long long>(8 * (kMaxPhases + 1));
  u.sched = c.take_n<int32_t>(kMaxSched);
  u.adam_tab = c.take_n<AdamTab>(1);
  u.fold_tab = c.take_n<FoldTab>(1);
  u.comm_tab = c.take_n<CommTab>(1);
  u.part = large ? nullptr : c.take_n<float>(int64_t(kSmallSplit) * kHidden * Bs);
  u.x_save = large ? nullptr : c.take_n<float>(Bs * kDin);
  u.x_stage = c.take_n<float>(Bs * (kDin + kDout));   // [x | y]: y starts right after the B x 32 block of x
  u.y_stage = u.x_stage ? u.x_stage + int64_t(B) * kDin : nullptr;
  u.p_stage = c.take_n<float>(Bs * kDout);
  u.scalars = c.take_n<float>(16);
  // accumulators cleared at the start of every step (one contiguous region)
  const int64_t acc0 = (c.off + 1023) & ~int64_t(1023);
  u.loss_acc = c.take_n<double>(16);
  u.gnorm2 = u.loss_acc ? u.loss_acc + 8 : nullptr;
  for (int l = 0; l < kNumHeavy; ++l) u.stats_f[l] = c.take_n<double>(2 * kHidden);
  for (int l = 0; l < kNumHeavy; ++l) u.stats_b[l] = c.take_n<double>(2 * kHidden);
  // chunk gates (helper-warp build): one counter per 2048 batch columns, per layer and direction; cleared with
  // the accumulators at the start of every step
  u.gate_words = ((Bs + kGateCols - 1) / kGateCols + 63) & ~int64_t(63);
  for (int l = 0; l < kNumHeavy; ++l) u.gates_f[l] = c.take_n<uint32_t>(u.gate_words);
  for (int l = 0; l < kNumHeavy; ++l) u.gates_b[l] = c.take_n<uint32_t>(u.gate_words);
  c.take(0);
  u.acc_base = c.base ? c.base + acc0 : nullptr;
  u.acc_bytes = ((c.off + 1023) & ~int64_t(1023)) - acc0;
  for (int l = 0; l < kNumHeavy; ++l) {
    u.save_mean[l] = c.take_n<float>(kHidden); u.save_rstd[l] = c.take_n<float>(kHidden);
    u.scale[l] = c.take_n<float>(kHidden); u.shift[l] = c.take_n<float>(kHidden);
  }
  // operand planes of the weights: ONE bf16 image of the flat parameter buffer per plane, so the Adam
  // phase writes them at the index it is already at; each weight matrix is TMA-mapped where it lies
  u.wplane_stride = (kPL.numel + 511) & ~int64_t(511);
  u.wplanes = c.take_n<uint16_t>(u.wplane_stride * P);
  for (int l = 0; l < kNumLinear; ++l) {
    u.W[l].p = u.wplanes ? u.wplanes + kPL.w[l] : nullptr;
    u.W[l].plane_stride = u.wplane_stride;
    u.W[l].ld = kPL.cols[l];
  }
  u.X_fm = take_planes(c, P, kDin, Bs);
  const int64_t zbytes = int64_t(kHidden) * Bs * (P == 1 ? 2 : 4);
  u.drop = nullptr;   // dropout bits are drawn where they are consumed (drop_word / drop_bits8)
  for (int l = 0; l < kNumHeavy; ++l) {
    u.h_fm[l] = take_planes(c, P, kHidden, Bs);
    if (train) {
      u.zT[l] = c.take(zbytes);
      u.dyT[l] = (large && l != 4 && l != 2) ? c.take(zbytes) : nullptr;   // layers 4 / 2 keep g instead
      u.keep[l] = c.take_n<uint32_t>(int64_t(kHidden) * ldw);
      u.dz_fm[l] = take_planes(c, P, kHidden, Bs);
    } else {
      u.zT[l] = u.dyT[l] = nullptr; u.keep[l] = nullptr;
      u.dz_fm[l] = PlaneRef{nullptr, 0, 0};
    }
  }
  if (train) {
    // gradients kept for the residual edges: fp32 on the small path, storage type of z on the large path
    const int64_t gbytes = int64_t(kHidden) * Bs * ((large && P == 1) ? 2 : 4);
    u.G2 = c.take(gbytes);
    u.G4 = c.take(gbytes);
    u.dP_fm = take_planes(c, P, kDout, Bs);
  } else {
    u.G2 = u.G4 = nullptr;
    u.dP_fm = PlaneRef{nullptr, 0, 0};
  }
  c.take(0);
  return c.off + 1024;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
struct Sched { int first = 0, count = 0; };
#ifndef B200_BLOCKED
#define B200_BLOCKED 0
#endif

struct b200_plan {
  b200_plan_desc d;
  Bufs u;
  int B, P, kind;
  bool large;
  int sms;
  int cg = 1;                // CTAs per MMA: 2 = CTA pairs (large-batch plans), 1 otherwise
  bool split_ew = false;     // large bf16 plans: the BatchNorm apply passes run as high-occupancy kernels of their own
  int tma_store = 0;         // large bf16 plans: epilogue rows leave through shared-memory staging tiles + TMA stores
                             // (bit 0: forward z, bit 1: dgrad g / dy)
  bool helpers = false;      // CTA pairs + helper warps: apply passes overlap the GEMM that consumes them (mega_pairx.cu)
  int variant() const { return helpers ? 3 : cg; }
  int workers() const { return sms / cg; }
  std::vector<CUtensorMap> tm;
  std::vector<PhaseBlob> blobs;
  std::vector<CtaTiles> cta_tiles;   // [phase][sms]
  std::vector<int32_t> sched;
  Sched s_refresh, s_eval, s_fwd, s_bwd, s_step, s_update, s_step_dp;
  bool has_comm = false;
  uint32_t comm_epoch = 1;   // flag value of the next cross-GPU barrier (flags start at 0)
  PhaseBlob cur;
  uint32_t bar_count = 0;   // grid-barrier arrivals so far (device counter is monotonic)
  int last_first = 0, last_count = 0;
};

static int add_map(b200_plan* pl, const PlaneRef& r, int64_t inner, int64_t rows, int box_rows, int* idx) {
  if (int(pl->tm.size()) >= kMaxTmaps) return fail("too many tensor maps");
  CUtensorMap m;
  if (encode_map(&m, r.p, inner, rows, pl->P, r.ld, r.plane_stride, box_rows)) return 1;
  pl->tm.push_back(m);
  *idx = int(pl->tm.size()) - 1;
  return 0;
}

static void begin_phase(b200_plan* pl) {
  memset(&pl->cur, 0, sizeof pl->cur);
  pl->cur.ph.aux_off = -1;
  pl->cur.ph.stage_bytes = 1024;
  pl->cur.ph.num_stages = 1;
  pl->cur.ph.sched_off = -1;
}
static int end_phase(b200_plan* pl) {
  Phase& p = pl->cur.ph;
  p.gemm_tiles = 0;
  int sb = 1024;
  for (int j = 0; j < p.num_gemm; ++j) {
    const GemmJob& g = pl->cur.gj[j];
    p.gemm_tiles += g.num_tiles;
    sb = std::max(sb, g.a.planes * (128 * 128 + (g.BN / pl->cg) * 128));
  }
  p.stage_bytes = sb;
  int ring = mega_ring_bytes();
  for (int j = 0; j < p.num_gemm; ++j)
    if (pl->cur.gj[j].a.tm_out) ring = mega_ring_bytes() - kEpiStageBytes;   // epilogue staging tiles sit at the top
  p.num_stages = std::max(1, std::min(8, ring / sb));
  p.aux_off = -1;
  if (pl->helpers && p.num_ew > 0) {
    // the apply pass keeps its per-feature constants in the part of the ring region the operand stages leave free
    if (mega_ring_bytes() - p.num_stages * sb < kAuxBytes && p.num_stages > 3) --p.num_stages;
    if (mega_ring_bytes() - p.num_stages * sb >= kAuxBytes) p.aux_off = p.num_stages * sb;
  }
  const int G = pl->workers();   // CTAs, or CTA pairs walking one tile list
  std::vector<CtaTiles> ct(G);
  for (int c = 0; c < G; ++c) memset(&ct[c], 0, sizeof(CtaTiles));
  if (p.gemm_tiles > 0) {
    // longest-processing-time-first assignment of tiles to CTAs (tiles of one phase differ in weight
    // when a wgrad slice shares the phase with dgrad tiles); ties keep index order, so CTAs working
    // at the same time read neighbouring tiles.
    struct T { int id; int64_t cost; };
    std::vector<T> tiles;
    tiles.reserve(p.gemm_tiles);
    int id = 0;
    for (int j = 0; j < p.num_gemm; ++j) {
      const GemmJob& g = pl->cur.gj[j];
      const int num_kb = (g.K + 63) / 64;
      const int pairs = g.a.planes == 1 ? 1 : 6;
      for (int t = 0; t < g.num_tiles; ++t, ++id) {
        const int ks = t / (g.m_tiles * g.n_tiles);
        const int kb = std::min(num_kb, (ks + 1) * g.kb_per_split) - ks * g.kb_per_split;
        // k-blocks x width, plus an epilogue term; the large-batch dgrad epilogue (z / keep / residual loads,
        // g or dy store, statistics) measures about as long as its 16 k-block mainloop
        const int64_t epi = (g.epi == EPI_DGRAD_LARGE) ? 16 * g.BN : 6 * g.BN;
        tiles.push_back(T{id, int64_t(kb) * pairs * g.BN + epi});
      }
    }
    bool gated = false;
    for (int j = 0; j < p.num_gemm; ++j) gated = gated || pl->cur.gj[j].gate != nullptr;
    if (gated) {
      // Tiles that wait for a chunk of the apply pass running in this phase: every CTA walks its list in CHUNK
      // order (the pass finishes chunk c before chunk c + 1), heavier tiles first within a chunk; tiles go to
      // the least-loaded CTA in that order, which balances the phase to within one tile.
      std::vector<int> chunk(tiles.size());
      int id2 = 0;
      for (int j = 0; j < p.num_gemm; ++j) {
        const GemmJob& g = pl->cur.gj[j];
        for (int t = 0; t < g.num_tiles; ++t, ++id2) {
          const int mt = t % g.m_tiles, nt = (t / g.m_tiles) % g.n_tiles, ks = t / (g.m_tiles * g.n_tiles);
          int hi = 0;
          if (g.gate == nullptr) hi = 0;
          else if (g.gate_kind == GATE_N) hi = std::min(g.N, (nt + 1) * g.BN) - 1;
          else if (g.gate_kind == GATE_K) hi = std::min(g.K, (ks + 1) * g.kb_per_split * 64) - 1;
          else hi = std::min(g.M, (mt + 1) * 128 * pl->cg) - 1;
          chunk[id2] = g.gate ? (hi >> kGateShift) : -1;   // ungated tiles of the phase can start at once
        }
      }
      std::stable_sort(tiles.begin(), tiles.end(), [&](const T& a, const T& b) {
        if (chunk[a.id] != chunk[b.id]) return chunk[a.id] < chunk[b.id];
        return a.cost > b.cost;
      });
    } else if (p.num_gemm == 2 && pl->large && !(getenv("B200_INTERLEAVE") && atoi(getenv("B200_INTERLEAVE")) == 0)) {
      // Two jobs in one phase = a dgrad GEMM (short mainloop, heavy epilogue: z / residual / keep rows in, dy or g
      // out) next to a wgrad GEMM (K slices with a mainloop twice as long and an epilogue that is a handful of vector
      // reductions).  Longest-first puts every wgrad slice in front and leaves each CTA a run of dgrad tiles whose
      // epilogues (13 us) cannot hide behind their own mainloops (4.3 us).  Interleaved by relative position --
      // about two dgrad tiles per wgrad slice on every CTA -- the dgrad epilogue of one tile runs under the wgrad
      // mainloop of the next (measured: 280 / 315 us -> 235 / 260 us per phase at batch 65 536).
      std::vector<double> key(tiles.size());
      int id2 = 0;
      for (int j = 0; j < p.num_gemm; ++j) {
        const GemmJob& g = pl->cur.gj[j];
        for (int t = 0; t < g.num_tiles; ++t, ++id2) key[id2] = (double(t) + 0.5) / double(g.num_tiles);
      }
      std::stable_sort(tiles.begin(), tiles.end(), [&](const T& a, const T& b) { return key[a.id] < key[b.id]; });
    } else {
      std::stable_sort(tiles.begin(), tiles.end(), [](const T& a, const T& b) { return a.cost > b.cost; });
    }
    std::vector<std::vector<int>> per(G);
    typedef std::pair<int64_t, int> LoadCta;
    std::priority_queue<LoadCta, std::vector<LoadCta>, std::greater<LoadCta>> pq;
    for (int c = 0; c < G; ++c) pq.push(LoadCta(0, c));
    for (const T& t : tiles) {
      LoadCta lc = pq.top();
      pq.pop();
      per[lc.second].push_back(t.id);
      pq.push(LoadCta(lc.first + t.cost, lc.second));
    }
    if (int64_t(pl->sched.size()) + p.gemm_tiles > kMaxSched) return fail("tile schedule table too small");
    p.sched_off = int(pl->sched.size());
    int off = 0;
    for (int c = 0; c < G; ++c) {
      ct[c].begin = off;
      off += int(per[c].size());
      ct[c].end = off;
      for (int i = 0; i < kTileSlots && i < int(per[c].size()); ++i) ct[c].tile[i] = per[c][i];
      pl->sched.insert(pl->sched.end(), per[c].begin(), per[c].end());
    }
  }
  if (int(pl->blobs.size()) >= kMaxPhases) return fail("too many phases");
  pl->blobs.push_back(pl->cur);
  pl->cta_tiles.insert(pl->cta_tiles.end(), ct.begin(), ct.end());
  return 0;
}

static Args base_args(const b200_plan* pl, int layer) {
  Args a;
  memset(&a, 0, sizeof a);
  a.layer = layer;
  a.planes = pl->P;
  a.B = pl->B;
  a.ldT = pl->u.Bs;
  a.ldw = pl->u.ldw;
  return a;
}

// D[m][n] = sum_k A[m][k] B[n][k].  K-major operand: global [rows = m|n][inner = k];
// MN-major operand (x_mn = 1): global [rows = k][inner = m|n] -- the feature-major activations.
static int add_gemm(b200_plan* pl, const PlaneRef& A, int a_mn, const PlaneRef& Bop, int b_mn, int M, int N, int K,
                    int BN, int epi, const Args& a, int ksplit = 1, int gate_kind = GATE_NONE, const uint32_t* gate = nullptr) {
  if (pl->cur.ph.num_gemm >= kBlobGemm) return fail("too many gemm jobs in one phase");
  if (b_mn && BN % 64) return fail("MN-major B operand needs a tile width that is a multiple of 64");
  GemmJob g;
  memset(&g, 0, sizeof g);
  if (a_mn ? add_map(pl, A, M, K, 64, &g.tmap_a) : add_map(pl, A, K, M, 128, &g.tmap_a)) return 1;
  const int cg = pl->cg;
  if (cg == 2 && ((BN / 2) % 8 || BN % 16 || (b_mn && BN % 128))) return fail("CTA-pair tile width must split into two halves");
  // CTA pairs: a tile is 256 rows of A (128 per CTA) and each CTA loads half of the B tile
  if (b_mn ? add_map(pl, Bop, N, K, 64, &g.tmap_b) : add_map(pl, Bop, K, N, BN / cg, &g.tmap_b)) return 1;
  g.a_mn = a_mn; g.b_mn = b_mn;
  g.M = M; g.N = N; g.K = K; g.BN = BN;
  g.m_tiles = (M + 128 * cg - 1) / (128 * cg);
  g.n_tiles = (N + BN - 1) / BN;
  const int num_kb = (K + 63) / 64;
  ksplit = std::max(1, std::min(ksplit, num_kb));
  g.kb_per_split = (num_kb + ksplit - 1) / ksplit;
  g.ksplit = (num_kb + g.kb_per_split - 1) / g.kb_per_split;   // no empty slices
  g.num_tiles = g.m_tiles * g.n_tiles * g.ksplit;
  g.epi = epi;
  g.gate_kind = gate ? gate_kind : GATE_NONE;
  g.gate = gate;
  g.a = a;
  g.a.tm_out = nullptr;
  if (a.planes == 1 && M % 128 == 0 &&
      ((epi == EPI_FWD_LARGE && (pl->tma_store & 1)) || (epi == EPI_DGRAD_LARGE && (pl->tma_store & 2)))) {
    // the epilogue's output rows leave through a staging tile and the TMA unit (Args::tm_out)
    void* dst = (epi == EPI_FWD_LARGE) ? a.zT : (a.g_out ? a.g_out : a.dyT);
    if (int(pl->tm.size()) >= kMaxTmaps) return fail("too many tensor maps");
    CUtensorMap m;
    if (encode_out_map(&m, dst, a.ldT, M, a.ldT)) return 1;
    pl->tm.push_back(m);
    g.a.tm_out = pl->u.tmaps + (pl->tm.size() - 1);
  }
  // large bf16 wgrad K-slices: the vector reductions leave through the epilogue warps' staging tiles (mega.cu, epi_wgrad);
  // tm_out is only a FLAG here (it also makes end_phase keep the staging region out of the operand ring)
  static const bool wgrad_stage_off = getenv("B200_WGRAD_STAGE") && atoi(getenv("B200_WGRAD_STAGE")) == 0;   // A/B switch
  if (epi == EPI_WGRAD && pl->large && a.planes == 1 && g.ksplit > 1 && M % (128 * cg) == 0 && N % 32 == 0 && !pl->helpers &&
      !wgrad_stage_off)
    g.a.tm_out = pl->u.tmaps;
  pl->cur.gj[pl->cur.ph.num_gemm++] = g;
  return 0;
}
static const GemmJob& last_gemm(const b200_plan* pl) { return pl->cur.gj[pl->cur.ph.num_gemm - 1]; }
static int add_ew(b200_plan* pl, int kind, const Args& a) {
  if (pl->cur.ph.num_ew >= kBlobEw) return fail("too many elementwise jobs in one phase");
  EwJob e;
  memset(&e, 0, sizeof e);
  e.kind = kind;
  e.a = a;
  pl->cur.ej[pl->cur.ph.num_ew++] = e;
  return 0;
}

static Args adam_args(const b200_plan* pl) {
  Args a = base_args(pl, 0);
  a.adam = pl->u.adam_tab;
  a.gnorm2 = pl->u.gnorm2;
  a.loss_acc = pl->u.loss_acc;
  return a;
}
static AdamTab make_adam_tab(const b200_plan* pl) {
  AdamTab t;
  memset(&t, 0, sizeof t);
  t.params = pl->d.params; t.grads = pl->d.grads; t.exp_avg = pl->d.exp_avg; t.exp_avg_sq = pl->d.exp_avg_sq;
  t.numel = kPL.numel;
  for (int l = 0; l < kNumLinear; ++l) { t.w_off[l] = kPL.w[l]; t.w_rows[l] = kPL.rows[l]; t.w_cols[l] = kPL.cols[l]; }
  t.planes = pl->u.wplanes;
  t.plane_stride = pl->u.wplane_stride;
  return t;
}
static FoldTab make_fold_tab(const b200_plan* pl) {
  FoldTab t;
  memset(&t, 0, sizeof t);
  for (int l = 0; l < kNumHeavy; ++l) {
    t.bias[l] = pl->d.params + kPL.b[l]; t.gamma[l] = pl->d.params + kPL.g[l]; t.beta[l] = pl->d.params + kPL.be[l];
    t.run_mean[l] = pl->d.running_mean[l]; t.run_var[l] = pl->d.running_var[l];
    t.scale[l] = pl->u.scale[l]; t.shift[l] = pl->u.shift[l];
  }
  return t;
}

// tile width over the batch dimension (activations are the MN-major B operand: multiples of 64)
static int tile_bn(const b200_plan* pl, int n) {
  const int cap = (pl->P == 1) ? 256 : 128;
  const int r = (n + 63) & ~63;
  return std::min(cap, r);
}
// tile width over a feature dimension (K-major B operand of the wgrad GEMMs)
static int tile_bn_feat(const b200_plan* pl, int n) {
  const int cap = (pl->P == 1) ? 256 : 128;
  const int r = (n + 15) & ~15;
  return std::min(cap, r);
}

// Args of the elementwise phase that follows a split-K GEMM on the small-batch path
static void small_part_args(const b200_plan* pl, Args& a) {
  a.out_f32 = pl->u.part;
  a.part = pl->u.part;
  a.ld_out = pl->u.Bs;
  a.part_stride = int64_t(kHidden) * pl->u.Bs;
}

// Args of forward heavy layer l (model/bilinear.py:7-13): shared by its GEMM epilogue and its apply pass
static Args fwd_layer_args(const b200_plan* pl, int l, bool train) {
  const Bufs& u = pl->u;
  const b200_plan_desc& d = pl->d;
  Args a = base_args(pl, l);
  a.bias = d.params + kPL.b[l];
  a.gamma = d.params + kPL.g[l];
  a.beta = d.params + kPL.be[l];
  a.run_mean = d.running_mean[l]; a.run_var = d.running_var[l]; a.nbt = d.num_batches_tracked[l];
  a.save_mean = u.save_mean[l]; a.save_rstd = u.save_rstd[l];
  a.scale = u.scale[l]; a.shift = u.shift[l];
  a.stats = u.stats_f[l];
  a.zT = u.zT[l];
  a.keep = u.keep[l];
  a.drop = u.drop ? u.drop + int64_t(l) * kHidden * u.ldw : nullptr;
  if (l == 2) a.skip_fm = u.h_fm[0];
  if (l == 4) a.skip_fm = u.h_fm[2];
  a.out_fm = u.h_fm[l];
  a.eval_mode = train ? 0 : 1;
  return a;
}

// forward heavy layer l (model/bilinear.py:7-13) -- train or eval flavour
static int build_fwd_layer(b200_plan* pl, int l, bool train, bool phase_open = false) {
  const Bufs& u = pl->u;
  const b200_plan_desc& d = pl->d;
  const int B = pl->B;
  Args a = fwd_layer_args(pl, l, train);
  const PlaneRef& in = (l == 0) ? u.X_fm : u.h_fm[l - 1];
  const int K = kPL.cols[l];
  const int BN = tile_bn(pl, B);
  if (pl->helpers && train) {
    // Helper-warp build: ONE phase per layer.  The apply pass of the PREVIOUS layer (z -> BatchNorm -> ReLU ->
    // dropout (+ skip) -> h, needs that layer's complete batch statistics, hence the grid barrier in front) runs
    // on the helper warps chunk by chunk while the GEMM roles multiply the chunks that are through.
    begin_phase(pl);
    pl->cur.ph.ew_first = 1;
    const uint32_t* gate = nullptr;
    if (l > 0) {
      Args pa = fwd_layer_args(pl, l - 1, true);
      pa.gate = u.gates_f[l - 1];
      gate = pa.gate;
      if (add_ew(pl, EW_BN_FWD_APPLY, pa)) return 1;
    }
    if (add_gemm(pl, u.W[l], 0, in, 1, kHidden, B, K, BN, EPI_FWD_LARGE, a, 1, GATE_N, gate)) return 1;
    return end_phase(pl);
  }
  if (!pl->large) {
    small_part_args(pl, a);
    a.drop = nullptr;   // dropout bits are drawn inline (drop_word)
    if (l == 0) {
      // the 32-wide first Linear runs on the CUDA cores inside the elementwise job; it joins the phase
      // the caller has open (accumulator clearing), so the forward pass starts with ONE phase
      a.part = nullptr; a.nsplit = 0;
      a.w_src = d.params + kPL.w[0];
      a.x_save = train ? u.x_save : nullptr;
      if (!phase_open) begin_phase(pl);
      if (add_ew(pl, EW_BN_FWD_SMALL, a)) return 1;
      if (end_phase(pl)) return 1;
      return 0;
    }
    begin_phase(pl);
    if (add_gemm(pl, u.W[l], 0, in, 1, kHidden, B, K, BN, EPI_PARTIAL, a, kSmallSplit)) return 1;
    a.nsplit = last_gemm(pl).ksplit;
    if (end_phase(pl)) return 1;
    begin_phase(pl);
    if (add_ew(pl, EW_BN_FWD_SMALL, a)) return 1;
    if (end_phase(pl)) return 1;
    return 0;
  }
  begin_phase(pl);
  if (add_gemm(pl, u.W[l], 0, in, 1, kHidden, B, K, BN, train ? EPI_FWD_LARGE : EPI_FWD_EVAL, a)) return 1;
  if (end_phase(pl)) return 1;
  if (train) {
    begin_phase(pl);
    if (add_ew(pl, EW_BN_FWD_APPLY, a)) return 1;
    if (end_phase(pl)) return 1;
  }
  return 0;
}

static int build_decode(b200_plan* pl, bool fuse_loss, bool train) {
  const Bufs& u = pl->u;
  Args a = base_args(pl, 5);
  a.bias = pl->d.params + kPL.b[5];
  a.fuse_loss = fuse_loss ? 1 : 0;
  a.loss_acc = u.loss_acc;
  a.grad_bias = pl->d.grads ? pl->d.grads + kPL.b[5] : nullptr;
  a.out_fm = u.dP_fm;
  if (!pl->large) {
    small_part_args(pl, a);
    if (train)
      for (int l = 0; l < kNumHeavy; ++l) a.nbt_w[l] = pl->d.num_batches_tracked[l];
    if (fuse_loss) a.gnorm2 = u.gnorm2;
    begin_phase(pl);
    if (add_gemm(pl, u.W[5], 0, u.h_fm[4], 1, kDout, pl->B, kHidden, tile_bn(pl, pl->B), EPI_PARTIAL, a, kSmallSplit))
      return 1;
    a.nsplit = last_gemm(pl).ksplit;
    if (end_phase(pl)) return 1;
    begin_phase(pl);
    if (add_ew(pl, EW_DECODE_SMALL, a)) return 1;
    if (end_phase(pl)) return 1;
    return 0;
  }
  begin_phase(pl);
  const uint32_t* gate = nullptr;
  if (pl->helpers && train) {
    // helper-warp build: the apply pass of the last heavy layer shares the phase with the decode GEMM
    pl->cur.ph.ew_first = 1;
    Args pa = fwd_layer_args(pl, kNumHeavy - 1, true);
    pa.gate = u.gates_f[kNumHeavy - 1];
    gate = pa.gate;
    if (add_ew(pl, EW_BN_FWD_APPLY, pa)) return 1;
  }
  {
    // Large batches: the decode GEMM runs TRANSPOSED -- P[b][48] = h[b][:] . W5^T with the activations as an
    // MN-major A operand (128 poses per tile) and the 48 x 1024 weight as the K-major B operand.  With rows =
    // output features only 48 of the 128 accumulator lanes (1.5 of the 4 epilogue warps) carry data and every
    // prediction / target access is a 4-byte access at a 192-byte stride; with rows = poses each lane owns
    // one pose: 192 contiguous bytes of target in, 192 contiguous bytes of prediction out.
    if (add_gemm(pl, u.h_fm[4], 1, u.W[5], 0, pl->B, kDout, kHidden, kDout, EPI_DECODE_T, a, 1, GATE_M, gate)) return 1;
  }
  if (end_phase(pl)) return 1;
  return 0;
}

// K slices of a large-batch wgrad GEMM (K = batch): slices of 32 k-blocks weigh as much as two dgrad
// mainloops, so the longest-first schedule balances the phase to within one such slice.  (Measured: 128
// k-block slices cut the L2 reduction traffic 4x but leave ~100 us of imbalance per phase, because the
// dgrad tiles are epilogue-bound and cost more than their k-blocks suggest -- see profiles/.)
static int wgrad_split(const b200_plan* pl) {
  if (!pl->large) return 1;
  const int num_kb = (pl->B + 63) / 64;
  static const int slice_kb = getenv("B200_WGRAD_SLICE") ? std::max(8, atoi(getenv("B200_WGRAD_SLICE"))) : 32;   // A/B switch
  return std::max(1, num_kb / slice_kb);
}

// backward phases (autograd of model/bilinear.py:31-41): for l = 4..0 { dgrad(l+1 -> l) with the
// BN/ReLU/dropout backward of layer l behind it, wgrad(l+1) } then wgrad(0).  The dgrad GEMMs read the
// [out][in] weight planes as an MN-major A operand (no transposed copy of the weights exists).
// `tail_ew`: elementwise job appended to the last (wgrad 0) phase, or nullptr.
static int build_backward(b200_plan* pl, bool with_norm, const Args* tail_ew, int tail_kind, bool prefetch_opt = false) {
  const Bufs& u = pl->u;
  const b200_plan_desc& d = pl->d;
  const int B = pl->B;
  const int wsplit = wgrad_split(pl);
  double* wnorm = (with_norm && wsplit == 1) ? u.gnorm2 : nullptr;   // split-K: EW_SUMSQ over everything instead
  static const bool thin_off = getenv("B200_THIN_STATS") && atoi(getenv("B200_THIN_STATS")) == 0;   // A/B switch
  auto layer_args = [&](int l) {
    Args a = base_args(pl, l);
    a.gamma = d.params + kPL.g[l];
    a.save_mean = u.save_mean[l]; a.save_rstd = u.save_rstd[l];
    a.zT = u.zT[l]; a.dyT = u.dyT[l];
    a.keep = u.keep[l];
    a.stats = u.stats_b[l];
    a.skip_in = (l == 2) ? u.G4 : (l == 0 ? u.G2 : nullptr);
    a.g_out = (l == 4) ? u.G4 : (l == 2 ? u.G2 : nullptr);
    a.out_fm = u.dz_fm[l];
    a.grad_bias = d.grads + kPL.b[l];
    a.grad_gamma = d.grads + kPL.g[l];
    a.grad_beta = d.grads + kPL.be[l];
    return a;
  };
  if (pl->helpers) {
    // Helper-warp build: the apply pass of layer l + 1 (dy, z -> dz; needs the complete sums of dy and dy * xhat)
    // shares a phase with the two GEMMs that consume dz_{l+1}: the dgrad towards layer l (dz is its B operand:
    // tiles gated by batch columns) and the wgrad of Linear l + 1 (dz is its K-major A operand: K slices gated).
    for (int l = 4; l >= -1; --l) {
      const int up = l + 1;
      begin_phase(pl);
      pl->cur.ph.ew_first = 1;
      const uint32_t* gate = nullptr;
      if (up <= 4) {
        Args pa = layer_args(up);
        pa.gate = u.gates_b[up];
        gate = pa.gate;
        if (add_ew(pl, EW_BN_BWD_APPLY, pa)) return 1;
      }
      const PlaneRef& dzu = (up == 5) ? u.dP_fm : u.dz_fm[up];
      if (l >= 0) {
        Args a = layer_args(l);
        if (add_gemm(pl, u.W[up], 1, dzu, 1, kHidden, B, kPL.rows[up], tile_bn(pl, B), EPI_DGRAD_LARGE, a, 1, GATE_N, gate))
          return 1;
      }
      Args w = base_args(pl, up);
      w.out_f32 = d.grads + kPL.w[up];
      w.ld_out = kPL.cols[up];
      w.gnorm2 = wnorm;
      const PlaneRef& hin = (l >= 0) ? u.h_fm[l] : u.X_fm;
      const int BNw = (l >= 0) ? tile_bn_feat(pl, kHidden) : 32;
      if (add_gemm(pl, dzu, 0, hin, 0, kPL.rows[up], kPL.cols[up], B, BNw, EPI_WGRAD, w, wsplit, GATE_K, gate)) return 1;
      if (end_phase(pl)) return 1;
    }
    (void)tail_ew; (void)tail_kind;
    return 0;
  }
  for (int l = 4; l >= 0; --l) {
    Args a = layer_args(l);
    const int up = l + 1;                      // the Linear whose input is h_l
    const int Kd = kPL.rows[up];               // 1024, or 48 for decode
    const PlaneRef& dzu = (up == 5) ? u.dP_fm : u.dz_fm[up];
    const bool thin = pl->split_ew && up == 5 && a.g_out != nullptr && !thin_off;
    begin_phase(pl);
    if (!pl->large && prefetch_opt) {
      pl->cur.ph.pf_slice = 4 - l;
      pl->cur.ph.pf_slices = 5;
    }
    if (!pl->large) {
      small_part_args(pl, a);
      if (add_gemm(pl, u.W[up], 1, dzu, 1, kHidden, B, Kd, tile_bn(pl, B), EPI_PARTIAL, a, kSmallSplit)) return 1;
      a.nsplit = last_gemm(pl).ksplit;
      a.gnorm2 = with_norm ? u.gnorm2 : nullptr;   // bias / BatchNorm-affine squares are added by the job itself
      if (l == 0) {
        // first layer: dW0 = dz0^T x is taken inside the elementwise job (K = 32), which also ends the
        // backward pass -> it publishes the loss when the optimizer phase does not follow
        a.grad_w0 = d.grads + kPL.w[0];
        a.x_save = u.x_save;
        a.out_fm = PlaneRef{nullptr, 0, 0};
        a.loss_acc = with_norm ? u.loss_acc : nullptr;
      }
    } else if (thin) {
      // the K = 48 dgrad stores the raw total gradient only (epi_dgrad_raw_bf16); EW_DGRAD_STATS follows
      Args r = a;
      r.stats = nullptr;
      if (add_gemm(pl, u.W[up], 1, dzu, 1, kHidden, B, Kd, tile_bn(pl, B), EPI_DGRAD_LARGE, r)) return 1;
    } else {
      if (add_gemm(pl, u.W[up], 1, dzu, 1, kHidden, B, Kd, tile_bn(pl, B), EPI_DGRAD_LARGE, a)) return 1;
    }
    {  // wgrad of Linear `up`: dW[out][in] = sum_b dz[b][out] h_l[b][in]
      Args w = base_args(pl, up);
      w.out_f32 = d.grads + kPL.w[up];
      w.ld_out = kHidden;
      w.gnorm2 = wnorm;
      if (add_gemm(pl, dzu, 0, u.h_fm[l], 0, kPL.rows[up], kHidden, B, tile_bn_feat(pl, kHidden), EPI_WGRAD, w, wsplit))
        return 1;
    }
    if (end_phase(pl)) return 1;
    if (thin) {
      begin_phase(pl);
      if (add_ew(pl, EW_DGRAD_STATS, a)) return 1;
      if (end_phase(pl)) return 1;
    }
    begin_phase(pl);
    if (add_ew(pl, pl->large ? EW_BN_BWD_APPLY : EW_BN_BWD_SMALL, a)) return 1;
    if (end_phase(pl)) return 1;
  }
  if (!pl->large) return 0;
  begin_phase(pl);
  {
    Args w = base_args(pl, 0);
    w.out_f32 = d.grads + kPL.w[0];
    w.ld_out = kDin;
    w.gnorm2 = wnorm;
    if (add_gemm(pl, u.dz_fm[0], 0, u.X_fm, 0, kHidden, kDin, B, 32, EPI_WGRAD, w, wsplit)) return 1;
  }
  if (tail_ew && add_ew(pl, tail_kind, *tail_ew)) return 1;
  if (end_phase(pl)) return 1;
  return 0;
}

static void zero_list_step(const b200_plan* pl, Args& a, bool fwd_acc, bool bwd_grads) {
  int k = 0;
  if (fwd_acc) { a.zero_ptr[k] = pl->u.acc_base; a.zero_bytes[k] = pl->u.acc_bytes; ++k; }
  if (bwd_grads) {
    if (pl->large) {
      // large batches: weight gradients are summed over K slices with reductions, heavy-layer bias
      // gradients with atomics -> the whole flat gradient starts from zero
      a.zero_ptr[k] = pl->d.grads; a.zero_bytes[k] = kPL.numel * 4; ++k;
    }
  }
}

static int build_plan(b200_plan* pl) {
  const Bufs& u = pl->u;
  const b200_plan_desc& d = pl->d;
  const bool train = pl->kind == B200_PLAN_TRAIN;

  // ---- refresh: master weights -> operand planes
  pl->s_refresh.first = int(pl->blobs.size());
  begin_phase(pl);
  if (add_ew(pl, EW_SPLIT_W, adam_args(pl))) return 1;
  if (end_phase(pl)) return 1;
  pl->s_refresh.count = int(pl->blobs.size()) - pl->s_refresh.first;

  // ---- eval forward (valid_bilinear.py:52)
  pl->s_eval.first = int(pl->blobs.size());
  if (pl->large) {
    begin_phase(pl);
    {
      Args a = base_args(pl, 0);
      a.fold = u.fold_tab;
      if (add_ew(pl, EW_BN_FOLD, a)) return 1;
    }
    {
      Args a = base_args(pl, 0);
      a.src = SRC_X; a.src_cols = kDin; a.out_fm = u.X_fm;
      if (add_ew(pl, EW_PREP, a)) return 1;
    }
    if (end_phase(pl)) return 1;
  }
  for (int l = 0; l < kNumHeavy; ++l)
    if (build_fwd_layer(pl, l, false)) return 1;
  if (build_decode(pl, false, false)) return 1;
  pl->s_eval.count = int(pl->blobs.size()) - pl->s_eval.first;
  if (!train) return 0;

  // first phase of a training-mode forward: clear the accumulators; large batches also split x into
  // operand planes, draw the dropout bits and bump the BatchNorm counters here (the small path does
  // those inline / at the end of the forward).  The small path leaves the phase OPEN for layer 0.
  auto prep_fwd = [&](bool zero_grads) -> int {
    begin_phase(pl);
    Args a = base_args(pl, 0);
    zero_list_step(pl, a, true, zero_grads);
    if (pl->large) {
      a.src = SRC_X; a.src_cols = kDin; a.out_fm = u.X_fm;
      for (int l = 0; l < kNumHeavy; ++l) a.nbt_w[l] = d.num_batches_tracked[l];
    }
    if (add_ew(pl, EW_PREP, a)) return 1;
    if (pl->large && end_phase(pl)) return 1;
    return 0;
  };

  // ---- training-mode forward alone (autograd path)
  pl->s_fwd.first = int(pl->blobs.size());
  if (prep_fwd(false)) return 1;
  for (int l = 0; l < kNumHeavy; ++l)
    if (build_fwd_layer(pl, l, true, l == 0 && !pl->large)) return 1;
  if (build_decode(pl, false, true)) return 1;
  pl->s_fwd.count = int(pl->blobs.size()) - pl->s_fwd.first;

  // ---- backward alone (autograd path): dP comes from the caller
  pl->s_bwd.first = int(pl->blobs.size());
  begin_phase(pl);
  {
    Args a = base_args(pl, 5);
    int k = 0;
    for (int l = 0; l < kNumHeavy; ++l) { a.zero_ptr[k] = u.stats_b[l]; a.zero_bytes[k] = 2 * kHidden * 8; ++k; }
    // the chunk gates of the backward apply passes lie contiguously behind each other (carve)
    a.zero_ptr[k] = u.gates_b[0]; a.zero_bytes[k] = (reinterpret_cast<uint8_t*>(u.gates_b[kNumHeavy - 1]) -
                                                     reinterpret_cast<uint8_t*>(u.gates_b[0])) + u.gate_words * 4; ++k;
    if (add_ew(pl, EW_PREP, a)) return 1;
    if (pl->large) {
      Args z = base_args(pl, 5);
      zero_list_step(pl, z, false, true);
      if (add_ew(pl, EW_PREP, z)) return 1;
    }
    Args s = base_args(pl, 5);
    s.out_fm = u.dP_fm; s.grad_bias = d.grads + kPL.b[5];
    if (add_ew(pl, EW_SPLIT_DP, s)) return 1;
  }
  if (end_phase(pl)) return 1;
  if (build_backward(pl, false, nullptr, 0)) return 1;
  pl->s_bwd.count = int(pl->blobs.size()) - pl->s_bwd.first;

  // ---- fused step (train_bilinear.py:75-83)
  pl->s_step.first = int(pl->blobs.size());
  if (prep_fwd(true)) return 1;
  for (int l = 0; l < kNumHeavy; ++l)
    if (build_fwd_layer(pl, l, true, l == 0 && !pl->large)) return 1;
  if (build_decode(pl, true, true)) return 1;
  if (!pl->large) {
    if (build_backward(pl, true, nullptr, 0, true)) return 1;
  } else {
    Args a = adam_args(pl);
    if (wgrad_split(pl) == 1 && !pl->helpers) {
      // the weight tiles were squared in the wgrad epilogues: only bias / BN-affine gradients are left,
      // and those were final before the last GEMM phase started
      a.src_cols = 1;
      if (build_backward(pl, true, &a, EW_SUMSQ)) return 1;
    } else {
      // (helper-warp build: the last apply pass shares the last GEMM phase, so the bias / BN-affine gradients are
      // final only after it -- the sum of squares always gets its own phase)
      a.src_cols = (wgrad_split(pl) == 1) ? 1 : 0;
      if (build_backward(pl, true, nullptr, 0)) return 1;
      begin_phase(pl);
      if (add_ew(pl, EW_SUMSQ, a)) return 1;
      if (end_phase(pl)) return 1;
    }
  }
  begin_phase(pl);
  if (add_ew(pl, EW_ADAM, adam_args(pl))) return 1;
  if (end_phase(pl)) return 1;
  pl->s_step.count = int(pl->blobs.size()) - pl->s_step.first;

  // ---- optimizer alone (data parallel: after the gradient all-reduce; autograd path)
  pl->s_update.first = int(pl->blobs.size());
  begin_phase(pl);
  {
    Args a = base_args(pl, 0);
    a.zero_ptr[0] = u.gnorm2; a.zero_bytes[0] = 64;
    if (add_ew(pl, EW_PREP, a)) return 1;
  }
  if (end_phase(pl)) return 1;
  begin_phase(pl);
  {
    Args a = adam_args(pl);
    a.src_cols = 0;
    a.loss_acc = nullptr;   // no forward ran in this launch: leave the caller's loss scalar alone
    if (add_ew(pl, EW_SUMSQ, a)) return 1;
  }
  if (end_phase(pl)) return 1;
  begin_phase(pl);
  {
    Args a = adam_args(pl);
    a.loss_acc = nullptr;
    if (add_ew(pl, EW_ADAM, a)) return 1;
  }
  if (end_phase(pl)) return 1;
  pl->s_update.count = int(pl->blobs.size()) - pl->s_update.first;

  // ---- fused DATA-PARALLEL step: forward, backward, in-kernel all-reduce of the flat gradient over NVLink
  // peer memory, clip + Adam on the sum (x 1/world) -- one persistent kernel per rank and iteration
  pl->s_step_dp.first = int(pl->blobs.size());
  if (prep_fwd(true)) return 1;
  for (int l = 0; l < kNumHeavy; ++l)
    if (build_fwd_layer(pl, l, true, l == 0 && !pl->large)) return 1;
  if (build_decode(pl, true, true)) return 1;
  if (build_backward(pl, false, nullptr, 0, true)) return 1;   // the clipping norm is taken on the REDUCED gradient
  for (int k = 0; k < 3; ++k) {
    Args a = base_args(pl, 0);
    a.comm = u.comm_tab;
    a.comm_step = (k == 2) ? 1 : 0;
    begin_phase(pl);
    if (add_ew(pl, k == 1 ? EW_ALLREDUCE : EW_COMM_BARRIER, a)) return 1;
    if (end_phase(pl)) return 1;
  }
  begin_phase(pl);
  {
    Args a = adam_args(pl);
    a.comm = u.comm_tab;
    if (add_ew(pl, EW_ADAM, a)) return 1;
  }
  if (end_phase(pl)) return 1;
  pl->s_step_dp.count = int(pl->blobs.size()) - pl->s_step_dp.first;
  return 0;
}

// One launch of the persistent kernel over phases [first, first + count) of the plan.
static int run_phases(b200_plan* pl, int first, int count, const Dyn& dyn, cudaStream_t stream) {
  MegaParams p;
  memset(&p, 0, sizeof p);
  p.blobs = pl->u.blobs + first;
  p.cta_tiles = pl->u.cta_tiles + size_t(first) * pl->workers();
  p.tmaps = pl->u.tmaps;
  p.sched = pl->u.sched;
  p.adam = (pl->d.exp_avg && pl->d.exp_avg_sq) ? pl->u.adam_tab : nullptr;
  p.barrier = pl->u.barrier;
  p.phase_ns = pl->u.phase_ns;
  pl->last_first = first; pl->last_count = count;
  p.num_phases = count;
  static const uint32_t dbg = getenv("B200_DBG") ? uint32_t(atoi(getenv("B200_DBG"))) : 0u;
  p.dbg = dbg;
  p.dyn = dyn;
  // B200_NOCOOP=1 (profiling only): ncu cannot launch the cooperative cluster kernel (LaunchFailed); the grid is one
  // CTA per SM, so all CTAs are co-resident in practice and the grid barrier still works (a 2 s timeout traps otherwise)
  static const bool nocoop = getenv("B200_NOCOOP") && atoi(getenv("B200_NOCOOP")) == 1;
  const bool coop = count > 1 && !nocoop;
  p.bar_base = pl->bar_count;
  const int grid = pl->workers() * pl->cg;
  pl->bar_count += uint32_t(grid) * uint32_t(count - 1);
  CK(launch_mega(p, grid, coop, stream, pl->variant()));
  g_launches.fetch_add(1);
  return 0;
}

// A phase that is nothing but one large-batch bf16 apply pass (classic builds): runs as a kernel of its own.
static bool is_split_apply(const b200_plan* pl, const PhaseBlob& b) {
  if (!pl->split_ew || b.ph.num_gemm != 0 || b.ph.num_ew != 1 || b.ej[0].a.planes != 1) return false;
  return b.ej[0].kind == EW_BN_FWD_APPLY || b.ej[0].kind == EW_BN_BWD_APPLY || b.ej[0].kind == EW_DGRAD_STATS;
}

static int run_sched(b200_plan* pl, const Sched& s, int count, const Dyn& dyn, cudaStream_t stream) {
  if (count <= 0) return fail("empty schedule (plan kind does not support this call)");
  // Stream order replaces the grid barrier on both sides of a split-out apply pass; the phases in between go to the
  // persistent kernel together (one cooperative launch per run of phases).
  int seg = s.first;
  for (int i = s.first; i < s.first + count; ++i) {
    const PhaseBlob& b = pl->blobs[i];
    if (!is_split_apply(pl, b)) continue;
    if (i > seg && run_phases(pl, seg, i - seg, dyn, stream)) return 1;
    if (b.ej[0].kind == EW_BN_FWD_APPLY) CK(launch_apply_fwd(b.ej[0].a, dyn, stream));
    else if (b.ej[0].kind == EW_DGRAD_STATS) CK(launch_dgrad_stats(b.ej[0].a, stream));
    else CK(launch_apply_bwd(b.ej[0].a, stream));
    g_launches.fetch_add(1);
    seg = i + 1;
  }
  if (s.first + count > seg && run_phases(pl, seg, s.first + count - seg, dyn, stream)) return 1;
  return 0;
}

static Dyn make_dyn(const b200_train_args* t, const b200_adam_args* a) {
  Dyn d;
  memset(&d, 0, sizeof d);
  for (int l = 0; l < kNumHeavy; ++l) d.momentum[l] = 0.1f;
  d.grad_scale = 1.f;
  d.bias_corr1 = 1.f; d.bias_corr2_sqrt = 1.f;
  if (t) {
    d.x = t->x; d.y = t->y; d.p_out = t->p_out; d.masks = t->masks;
    d.seed = t->seed; d.step = t->step;
    d.loss_out = t->loss_out; d.gnorm_out = t->gnorm_out;
    for (int l = 0; l < kNumHeavy; ++l) d.momentum[l] = t->momentum[l];
  }
  if (a) {
    d.lr = a->lr; d.beta1 = a->beta1; d.beta2 = a->beta2; d.adam_eps = a->eps;
    // torch/optim/adam.py: bias_correction1 = 1 - beta1**step, bias_correction2_sqrt = sqrt(1 - beta2**step)
    d.bias_corr1 = float(1.0 - std::pow(double(a->beta1), double(a->step)));
    d.bias_corr2_sqrt = float(std::sqrt(1.0 - std::pow(double(a->beta2), double(a->step))));
    d.max_norm = a->max_norm;
    d.grad_scale = a->grad_scale;
  }
  return d;
}

__global__ void unblock_keep_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, int ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kHidden * ldw) return;
  const int f = i / ldw, w = i - f * ldw;
#if B200_BLOCKED
  dst[i] = src[((int64_t(f >> 5) * ldw + w) << 5) + (f & 31)];
#else
  dst[i] = src[i];
#endif
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int b200_version(void) { return 100; }
const char* b200_last_error(void) { return g_err.c_str(); }
int64_t b200_param_numel(void) { return kPL.numel; }
int64_t b200_launch_count(void) { return g_launches.load(); }

int64_t b200_workspace_bytes(int32_t batch, int32_t precision, int32_t kind) {
  if (batch < 1 || (precision != 1 && precision != 3)) return -1;
  Carver c(nullptr);
  Bufs u;
  return carve(c, u, batch, precision, kind);
}

int b200_plan_create(const b200_plan_desc* desc, b200_plan** out) {
  if (!desc || !out) return fail("null argument");
  if (desc->batch < 1) return fail("batch must be >= 1");
  if (desc->precision != B200_PREC_BF16 && desc->precision != B200_PREC_FP32) return fail("bad precision");
  if (desc->kind != B200_PLAN_TRAIN && desc->kind != B200_PLAN_INFER) return fail("bad plan kind");
  if (!desc->params || !desc->workspace) return fail("params and workspace are required");
  if (desc->kind == B200_PLAN_TRAIN && !desc->grads) return fail("training plan needs a gradient buffer");
  if (reinterpret_cast<uintptr_t>(desc->workspace) % 1024) return fail("workspace must be 1024-byte aligned");
  const int64_t need = b200_workspace_bytes(desc->batch, desc->precision, desc->kind);
  if (desc->workspace_bytes < need)
    return fail("workspace too small: %lld < %lld", (long long)desc->workspace_bytes, (long long)need);
  for (int l = 0; l < kNumHeavy; ++l)
    if (!desc->running_mean[l] || !desc->running_var[l]) return fail("BatchNorm buffers are required");
  b200_plan* pl = new b200_plan();
  pl->d = *desc;
  pl->B = desc->batch; pl->P = desc->precision; pl->kind = desc->kind;
  pl->large = pl->B > (pl->P == 1 ? 256 : 128);
  // Large-batch plans run on CTA pairs (tcgen05 cta_group::2: clusters of two CTAs share every MMA, M = 256, each
  // CTA loads its 128 rows of A and HALF of the B tile): 32 KB instead of 48 KB per k-block go through each SM's
  // shared memory, whose bandwidth -- TMA fill + tensor-core operand reads -- bounds the single-CTA 128 x 256 tile
  // (DESIGN.md section 8).  B200_CG=1 keeps single-CTA MMAs (A/B comparison, parity-tested as well).
  pl->cg = (pl->large && !(getenv("B200_CG") && atoi(getenv("B200_CG")) == 1)) ? 2 : 1;
  int dev = 0;
  cudaGetDevice(&dev);
  pl->sms = 148;
  cudaDeviceGetAttribute(&pl->sms, cudaDevAttrMultiProcessorCount, dev);
  // pairs need one resident cluster per two SMs under the cooperative launch; otherwise the single-CTA build
  if (pl->cg == 2 && mega_pair_max_clusters(pl->sms) < pl->sms / 2) pl->cg = 1;
  // B200_HELPERS=1 (experiment, off by default): CTA pairs + helper warps (24 warps per CTA), the BatchNorm apply passes
  // of a training plan overlap the GEMM that consumes their output (mega_pairx.cu; measured in DESIGN.md section 8).
  pl->helpers = pl->cg == 2 && getenv("B200_HELPERS") && atoi(getenv("B200_HELPERS")) == 1 &&
                mega_pairx_max_clusters(pl->sms) >= pl->sms / 2;
  // B200_SPLIT_EW=0 keeps the apply passes inside the persistent kernel (per-phase time stamps, A/B comparison)
  pl->split_ew = pl->large && pl->P == 1 && !pl->helpers && !(getenv("B200_SPLIT_EW") && atoi(getenv("B200_SPLIT_EW")) == 0);
  // B200_TMA_STORE=<mask> (experiment, off by default).  Measured at batch 65536: inside one persistent launch the
  // forward GEMM phases lose 4.5 us each with staged TMA stores of z and the dgrad phases GAIN 3 .. 14 us with staged
  // stores of g / dy; with the apply passes split out (the default) the step time does not move either way, and the
  // epilogue-bound thin phases (K = 32 / 48) stay as long as they were -- their epilogues do not wait for the
  // load/store unit.
  pl->tma_store = (B200_EPI_TMA_HOST && pl->large && pl->P == 1 && !pl->helpers && !B200_BLOCKED && getenv("B200_TMA_STORE"))
                      ? (atoi(getenv("B200_TMA_STORE")) & 3) : 0;
  Carver c(desc->workspace);
  carve(c, pl->u, pl->B, pl->P, pl->kind);
  if (build_plan(pl)) { delete pl; return 1; }
  // schedules + tensor maps -> device (pageable host memory: cudaMemcpy is synchronous, done once)
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaMemcpy(pl->u.tmaps, pl->tm.data(), pl->tm.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(pl->u.blobs, pl->blobs.data(), pl->blobs.size() * sizeof(PhaseBlob), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaMemcpy(pl->u.cta_tiles, pl->cta_tiles.data(), pl->cta_tiles.size() * sizeof(CtaTiles), cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !pl->sched.empty())
    e = cudaMemcpy(pl->u.sched, pl->sched.data(), pl->sched.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    const AdamTab tab = make_adam_tab(pl);
    e = cudaMemcpy(pl->u.adam_tab, &tab, sizeof tab, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) {
    const FoldTab tab = make_fold_tab(pl);
    e = cudaMemcpy(pl->u.fold_tab, &tab, sizeof tab, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemset(pl->u.barrier, 0, 256);
  if (e != cudaSuccess) { delete pl; return fail("plan upload failed: %s", cudaGetErrorString(e)); }
  *out = pl;
  return 0;
}

void b200_plan_destroy(b200_plan* plan) { delete plan; }

int b200_refresh_weights(b200_plan* plan, void* stream) {
  if (!plan) return fail("null plan");
  Dyn d = make_dyn(nullptr, nullptr);
  return run_sched(plan, plan->s_refresh, plan->s_refresh.count, d, (cudaStream_t)stream);
}

int b200_forward_eval(b200_plan* plan, const float* x, float* p, void* stream) {
  if (!plan || !x || !p) return fail("null argument");
  Dyn d = make_dyn(nullptr, nullptr);
  d.x = x; d.p_out = p;
  return run_sched(plan, plan->s_eval, plan->s_eval.count, d, (cudaStream_t)stream);
}

int b200_forward_train(b200_plan* plan, const b200_train_args* args, void* stream) {
  if (!plan || !args || !args->x) return fail("null argument");
  if (plan->B < 2) return fail("Expected more than 1 value per channel when training, got input size [1, 1024]");
  Dyn d = make_dyn(args, nullptr);
  d.y = nullptr; d.loss_out = nullptr; d.gnorm_out = nullptr;
  return run_sched(plan, plan->s_fwd, plan->s_fwd.count, d, (cudaStream_t)stream);
}

int b200_backward(b200_plan* plan, const float* dP, void* stream) {
  if (!plan || !dP) return fail("null argument");
  Dyn d = make_dyn(nullptr, nullptr);
  d.dP = dP;
  return run_sched(plan, plan->s_bwd, plan->s_bwd.count, d, (cudaStream_t)stream);
}

int b200_train_step(b200_plan* plan, const b200_train_args* targs, const b200_adam_args* aargs, int32_t update,
                    void* stream) {
  if (!plan || !targs || !targs->x || !targs->y) return fail("null argument");
  if (update && (!aargs || !plan->d.exp_avg || !plan->d.exp_avg_sq)) return fail("optimizer state / args missing");
  if (plan->B < 2) return fail("Expected more than 1 value per channel when training, got input size [1, 1024]");
  Dyn d = make_dyn(targs, update ? aargs : nullptr);
  // the mean loss is produced by the last phase that runs: Adam when updating, otherwise a tiny tail
  int count = plan->s_step.count - (update ? 0 : 1);
  int rc = run_sched(plan, plan->s_step, count, d, (cudaStream_t)stream);
  if (rc) return rc;
  return 0;
}

// ---- data parallel over NVLink peer memory --------------------------------------------------------
int b200_comm_alloc(int64_t bytes, void** dev_ptr, void* handle64) {
  if (!dev_ptr || !handle64 || bytes <= 0) return fail("bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  CK(cudaMalloc(&p, size_t(bytes)));
  CK(cudaMemset(p, 0, size_t(bytes)));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return fail("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e)); }
  memcpy(handle64, &h, 64);
  *dev_ptr = p;
  return 0;
}
int b200_comm_open(const void* handle64, void** peer_ptr) {
  if (!handle64 || !peer_ptr) return fail("null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  CK(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int b200_comm_close(void* peer_ptr) {
  if (peer_ptr) CK(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}
int b200_comm_free(void* dev_ptr) {
  if (dev_ptr) CK(cudaFree(dev_ptr));
  return 0;
}
int64_t b200_comm_bytes(void) { return int64_t(kPL.numel) * 4 + int64_t(sizeof(CommCtl)); }

int b200_plan_set_comm(b200_plan* plan, const b200_comm_desc* c) {
  if (!plan || !c) return fail("null argument");
  if (c->world < 1 || c->world > kMaxRanks || c->rank < 0 || c->rank >= c->world) return fail("bad rank / world size");
  if (plan->kind != B200_PLAN_TRAIN) return fail("training plan required");
  if (c->buffers[c->rank] != (void*)plan->d.grads)
    return fail("the plan's gradient buffer must be this rank's communication buffer");
  CommTab t;
  memset(&t, 0, sizeof t);
  for (int r = 0; r < c->world; ++r) {
    if (!c->buffers[r]) return fail("missing peer buffer");
    t.grads[r] = reinterpret_cast<float*>(c->buffers[r]);
    t.ctl[r] = reinterpret_cast<CommCtl*>(reinterpret_cast<uint8_t*>(c->buffers[r]) + int64_t(kPL.numel) * 4);
  }
  t.rank = c->rank; t.world = c->world; t.numel = kPL.numel;
  t.sq_local = plan->u.loss_acc + 12;
  CK(cudaMemcpy(plan->u.comm_tab, &t, sizeof t, cudaMemcpyHostToDevice));
  plan->has_comm = true;
  plan->comm_epoch = c->epoch ? c->epoch : 1;
  return 0;
}

int b200_train_step_dp(b200_plan* plan, const b200_train_args* targs, const b200_adam_args* aargs, uint32_t* epoch_io,
                       void* stream) {
  if (!plan || !targs || !targs->x || !targs->y || !aargs) return fail("null argument");
  if (!plan->has_comm) return fail("b200_plan_set_comm has not been called on this plan");
  if (!plan->d.exp_avg || !plan->d.exp_avg_sq) return fail("optimizer state missing");
  if (plan->B < 2) return fail("Expected more than 1 value per channel when training, got input size [1, 1024]");
  Dyn d = make_dyn(targs, aargs);
  // the flags are shared by all plans of this process: the caller carries the epoch across plans
  if (epoch_io && *epoch_io) plan->comm_epoch = *epoch_io;
  d.comm_epoch = plan->comm_epoch;
  plan->comm_epoch += 2;
  if (epoch_io) *epoch_io = plan->comm_epoch;
  return run_sched(plan, plan->s_step_dp, plan->s_step_dp.count, d, (cudaStream_t)stream);
}

int b200_train_step_host(b200_plan* plan, const float* x_host, const float* y_host, const b200_train_args* targs,
                         const b200_adam_args* aargs, int32_t update, float* out_host, void* stream) {
  if (!plan || !x_host || !y_host || !targs || !out_host) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = size_t(plan->B);
  // x and y live next to each other in the workspace: when the caller's batches are adjacent in host
  // memory too (one pinned [x | y] block) they travel as ONE copy
  bool one_copy = false;
  if (y_host == x_host + B * kDin) {
    // adjacent addresses may still be two separate pinned allocations: then the copy is rejected up front
    // (synchronously, nothing enqueued) and the batches go one by one
    one_copy = cudaMemcpyAsync(plan->u.x_stage, x_host, B * (kDin + kDout) * sizeof(float), cudaMemcpyHostToDevice, st) ==
               cudaSuccess;
    if (!one_copy) cudaGetLastError();
  }
  if (!one_copy) {
    CK(cudaMemcpyAsync(plan->u.x_stage, x_host, B * kDin * sizeof(float), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(plan->u.y_stage, y_host, B * kDout * sizeof(float), cudaMemcpyHostToDevice, st));
  }
  b200_train_args t = *targs;
  t.x = plan->u.x_stage; t.y = plan->u.y_stage;
  // the two result scalars are stored straight into the caller's host memory when the device can address
  // it (pinned memory under unified addressing): no device -> host copy operation at all
  float* dev_view = nullptr;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, out_host) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer)
    dev_view = reinterpret_cast<float*>(attr.devicePointer);
  else
    cudaGetLastError();   // pageable memory: not an error, fall back to a copy
  t.loss_out = dev_view ? dev_view : plan->u.scalars;
  t.gnorm_out = dev_view ? dev_view + 1 : plan->u.scalars + 1;
  int rc = b200_train_step(plan, &t, aargs, update, stream);
  if (rc) return rc;
  if (!dev_view) CK(cudaMemcpyAsync(out_host, plan->u.scalars, 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

int b200_forward_eval_host(b200_plan* plan, const float* x_host, float* p_host, void* stream) {
  if (!plan || !x_host || !p_host) return fail("null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t B = size_t(plan->B);
  CK(cudaMemcpyAsync(plan->u.x_stage, x_host, B * kDin * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = b200_forward_eval(plan, plan->u.x_stage, plan->u.p_stage, stream);
  if (rc) return rc;
  CK(cudaMemcpyAsync(p_host, plan->u.p_stage, B * kDout * sizeof(float), cudaMemcpyDeviceToHost, st));
  return 0;
}

int b200_adam_step(b200_plan* plan, const b200_adam_args* aargs, float* gnorm_out, void* stream) {
  if (!plan || !aargs) return fail("null argument");
  if (!plan->d.exp_avg || !plan->d.exp_avg_sq) return fail("optimizer state missing");
  Dyn d = make_dyn(nullptr, aargs);
  d.gnorm_out = gnorm_out;
  return run_sched(plan, plan->s_update, plan->s_update.count, d, (cudaStream_t)stream);
}

static int eval_common(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n,
                       int32_t joints, const int32_t* action, int32_t n_actions, float* dist_out, float* pdist_out,
                       double* acc_sum, double* pacc_sum, int64_t* acc_cnt, bool do_m, bool do_p, void* stream) {
  if (!pred || !gt) return fail("null argument");
  if (n < 0) return fail("negative pose count");
  if (joints != 16 && joints != 17) return fail("joints must be 16 or 17");
  if ((mean == nullptr) != (stddev == nullptr)) return fail("mean and stddev must be given together");
  if (n_actions < 1 || n_actions > 64) return fail("n_actions must be in [1, 64]");
  if (n == 0) return 0;
  EvalParams prm;
  prm.pred = pred; prm.gt = gt; prm.mean = mean; prm.stddev = stddev; prm.action = action;
  prm.dist_out = dist_out; prm.pdist_out = pdist_out; prm.acc_sum = acc_sum; prm.pacc_sum = pacc_sum;
  prm.acc_cnt = reinterpret_cast<long long*>(acc_cnt);
  prm.n = n; prm.n_actions = n_actions;
  // J = 16 with 16-byte aligned tensors: TMA-fed kernel ([n][48] fp32 viewed through two tensor maps
  // that travel in the kernel parameters); otherwise (J = 17: 204-byte rows) the register-staged one
  EncodeTiledFn enc = get_encode();
  if (joints == 16 && enc && n < (int64_t(1) << 31) && reinterpret_cast<uintptr_t>(pred) % 16 == 0 &&
      reinterpret_cast<uintptr_t>(gt) % 16 == 0) {
    CUtensorMap tm[2];
    const void* bases[2] = {pred, gt};
    for (int i = 0; i < 2; ++i) {
      cuuint64_t dims[2] = {48, cuuint64_t(n)};
      cuuint64_t strides[1] = {192};
      cuuint32_t box[2] = {16, cuuint32_t(eval_tma_box_rows())};
      cuuint32_t estr[2] = {1, 1};
      CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(bases[i]), dims, strides, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled (eval) -> %d", int(r));
    }
    int launches = 1;
    CK(launch_eval_tma(prm, &tm[0], &tm[1], do_m, do_p, (cudaStream_t)stream, &launches));
    g_launches.fetch_add(launches - 1);
  } else {
    CK(launch_eval(prm, joints, do_m, do_p, (cudaStream_t)stream));
  }
  g_launches.fetch_add(1);
  return 0;
}

int b200_mpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n, int32_t joints,
               const int32_t* action, int32_t n_actions, float* dist_out, double* acc_sum, int64_t* acc_cnt,
               void* stream) {
  return eval_common(pred, gt, mean, stddev, n, joints, action, n_actions, dist_out, nullptr, acc_sum, nullptr, acc_cnt,
                     true, false, stream);
}
int b200_pmpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n, int32_t joints,
                const int32_t* action, int32_t n_actions, float* dist_out, double* acc_sum, int64_t* acc_cnt,
                void* stream) {
  return eval_common(pred, gt, mean, stddev, n, joints, action, n_actions, nullptr, dist_out, nullptr, acc_sum, acc_cnt,
                     false, true, stream);
}
int b200_mpjpe_pmpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n,
                      int32_t joints, const int32_t* action, int32_t n_actions, float* dist_out, float* pdist_out,
                      double* acc_sum, double* pacc_sum, int64_t* acc_cnt, void* stream) {
  return eval_common(pred, gt, mean, stddev, n, joints, action, n_actions, dist_out, pdist_out, acc_sum, pacc_sum,
                     acc_cnt, true, true, stream);
}

// ReLU/dropout decisions of the last training-mode forward (tests: admit either decision for elements
// within round-off of the ReLU kink).  out: [1024][ceil(B/32)] words, bit j of word w <-> row 32 w + j.
int b200_debug_keep_bits(b200_plan* plan, int32_t layer, uint32_t* out, void* stream) {
  if (!plan || !out) return fail("null argument");
  if (layer < 0 || layer >= kNumHeavy) return fail("layer must be in [0, 5)");
  if (plan->kind != B200_PLAN_TRAIN || !plan->u.keep[layer]) return fail("training plan required");
  if (plan->large) {
    // large-batch plans keep the words in 32-feature bands (mega.cu: blk_word): hand them out as [1024][ldw]
    const int total = kHidden * int(plan->u.ldw);
    unblock_keep_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(plan->u.keep[layer], out, int(plan->u.ldw));
    CK(cudaGetLastError());
    return 0;
  }
  CK(cudaMemcpyAsync(out, plan->u.keep[layer], size_t(kHidden) * plan->u.ldw * 4, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return 0;
}

// Per-phase durations (ns) of the most recent launch of this plan; synchronises the device.
// kinds_out[i]: 0 = GEMM phase (value = number of tiles in tiles_out[i]), else 1 + EwKind of its first job.
int b200_debug_phase_ns(b200_plan* plan, int64_t* ns_out, int32_t* kinds_out, int32_t* tiles_out, int32_t max_phases,
                        int32_t* count_out) {
  if (!plan || !ns_out || !count_out) return fail("null argument");
  CK(cudaDeviceSynchronize());
  std::vector<unsigned long long> t(8 * (plan->last_count + 1));
  CK(cudaMemcpy(t.data(), plan->u.phase_ns, t.size() * 8, cudaMemcpyDeviceToHost));
  const int n = std::min(plan->last_count, max_phases);
  for (int i = 0; i < n; ++i) {
    // 8 values per phase: total duration, then offsets from the phase start of: producer start,
    // producer done, MMA issue done, first accumulator ready, epilogue done, work done (block 0), spare
    ns_out[8 * i] = int64_t(t[8 * (i + 1)] - t[8 * i]);
    for (int k = 1; k < 8; ++k) ns_out[8 * i + k] = t[8 * i + k] >= t[8 * i] ? int64_t(t[8 * i + k] - t[8 * i]) : -1;
    if (i == 0) ns_out[7] = int64_t(t[0] - t[7]);   // kernel entry -> start of the first phase
    const PhaseBlob& bl = plan->blobs[plan->last_first + i];
    const Phase& ph = bl.ph;
    if (kinds_out) kinds_out[i] = ph.gemm_tiles > 0 ? 0 : 1 + bl.ej[0].kind;
    if (tiles_out) tiles_out[i] = ph.gemm_tiles > 0 ? (ph.gemm_tiles | (bl.gj[0].epi << 24)) : ph.num_ew;
  }
  *count_out = n;
  return 0;
}

// ---- debug GEMM: one single-phase schedule built on the fly in `scratch` (>= 64 KiB)
int b200_debug_gemm(const uint16_t* a, const uint16_t* b, float* dmat, int32_t m, int32_t n, int32_t k, int32_t planes,
                    int32_t a_mn, int32_t b_mn, int32_t bn, void* scratch, int64_t scratch_bytes, void* stream) {
  if (!a || !b || !dmat || !scratch) return fail("null argument");
  if (scratch_bytes < 65536 || reinterpret_cast<uintptr_t>(scratch) % 1024) return fail("scratch: 64 KiB, 1 KiB aligned");
  if (bn % 16 || bn < 16 || bn > 256 || (b_mn && bn % 64) || (planes == 3 && bn > 128)) return fail("bad bn");
  if (planes != 1 && planes != 3) return fail("bad planes");
  uint8_t* s = reinterpret_cast<uint8_t*>(scratch);
  CUtensorMap tm[2];
  if (!a_mn) { if (encode_map(&tm[0], a, k, m, planes, k, int64_t(m) * k, 128)) return 1; }
  else       { if (encode_map(&tm[0], a, m, k, planes, m, int64_t(m) * k, 64)) return 1; }
  if (!b_mn) { if (encode_map(&tm[1], b, k, n, planes, k, int64_t(n) * k, bn)) return 1; }
  else       { if (encode_map(&tm[1], b, n, k, planes, n, int64_t(n) * k, 64)) return 1; }
  GemmJob g;
  memset(&g, 0, sizeof g);
  g.tmap_a = 0; g.tmap_b = 1; g.a_mn = a_mn; g.b_mn = b_mn;
  g.M = m; g.N = n; g.K = k; g.BN = bn;
  g.m_tiles = (m + 127) / 128; g.n_tiles = (n + bn - 1) / bn; g.num_tiles = g.m_tiles * g.n_tiles;
  g.ksplit = 1; g.kb_per_split = (k + 63) / 64;
  g.epi = EPI_STORE_F32;
  g.a.planes = planes; g.a.out_f32 = dmat; g.a.ld_out = n;
  static PhaseBlob blob;   // host staging (one debug call at a time)
  memset(&blob, 0, sizeof blob);
  blob.ph.num_gemm = 1; blob.ph.gemm_tiles = g.num_tiles; blob.ph.sched_off = -1;
  blob.ph.stage_bytes = planes * (128 * 128 + bn * 128);
  blob.ph.num_stages = std::max(1, std::min(8, mega_ring_bytes() / blob.ph.stage_bytes));
  blob.gj[0] = g;
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(s, tm, sizeof tm, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(s + 1024, &blob, sizeof blob, cudaMemcpyHostToDevice, st));
  CK(cudaStreamSynchronize(st));  // the host staging copies above must be consumed before we return
  MegaParams p;
  memset(&p, 0, sizeof p);
  p.tmaps = s; p.blobs = reinterpret_cast<PhaseBlob*>(s + 1024);
  p.barrier = reinterpret_cast<unsigned int*>(s + 32768); p.num_phases = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  CK(launch_mega(p, std::min(sms, g.num_tiles), false, st, 1));
  g_launches.fetch_add(1);
  return 0;
}

__global__ void split_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int64_t numel, int planes) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < numel; i += int64_t(gridDim.x) * blockDim.x) {
    float v = src[i];
    for (int p = 0; p < planes; ++p) {
      __nv_bfloat16 h = __float2bfloat16_rn(v);
      dst[int64_t(p) * numel + i] = __bfloat16_as_ushort(h);
      v -= __bfloat162float(h);
    }
  }
}
int b200_debug_split(const float* src, uint16_t* dst, int64_t numel, int32_t planes, void* stream) {
  if (!src || !dst) return fail("null argument");
  const int grid = int(std::min<int64_t>(4096, (numel + 255) / 256));
  split_kernel<<<std::max(grid, 1), 256, 0, (cudaStream_t)stream>>>(src, dst, numel, planes);
  CK(cudaGetLastError());
  g_launches.fetch_add(1);
  return 0;
}

}  // extern "C"
