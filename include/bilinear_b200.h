/*
 * bilinear_b200 -- C ABI of the B200 (sm_100a) implementation of the nulledge/bilinear 2D->3D
 * pose-lifting hot path.
 *
 * The reference has no FFI layer: its hot path sits behind a Python nn.Module
 * (model/bilinear.py:16-41 BilinearUnit), a factory (model/bilinear.py:58-92 load) and two scripts
 * (train_bilinear.py:56-90, valid_bilinear.py:36-83), and all arithmetic is executed by PyTorch
 * (nn.Linear / nn.BatchNorm1d / nn.ReLU / nn.Dropout / nn.MSELoss / clip_grad_norm_ / optim.Adam).
 * This library is what a maintainer binds UNDER that Python interface (ctypes stub: see
 * INTEGRATION.md and bilinear_b200/_lib.py).  Each entry point names the reference call it replaces.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; b200_last_error() gives the message
 *     (thread-local, valid until the next failing call on that thread);
 *   - all pointers are DEVICE pointers unless the name ends in _host; the library never allocates or
 *     frees caller-visible memory: parameters, optimizer state, activations workspace are passed in;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); calls are asynchronous
 *     with respect to the host, ordered on that stream, and must come from one host thread per plan;
 *   - tensors are fp32, row-major, contiguous, exactly as the reference's tensors:
 *       x  [B][32]   (2 * 16 joints, model/bilinear.py:20-22)
 *       y,p[B][48]   (3 * 16 joints, model/bilinear.py:29)
 *     parameters are ONE flat fp32 buffer in the order of BilinearUnit.parameters()
 *     (the same order torch.optim.Adam indexes them, model/bilinear.py:60):
 *       encode.0.weight[1024][32] encode.0.bias[1024] encode.1.weight[1024] encode.1.bias[1024]
 *       bilinear.{0,1}.{0,1}.0.weight[1024][1024] .0.bias[1024] .1.weight[1024] .1.bias[1024]  (x4)
 *       decode.weight[48][1024] decode.bias[48]                         -> 4 291 632 floats.
 */
#ifndef BILINEAR_B200_H_
#define BILINEAR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_PARAM_NUMEL 4291632
#define B200_NUM_BN 5

/* precision of the GEMM operands */
#define B200_PREC_BF16 1 /* bf16 operands, fp32 accumulate (tolerance 2e-2 rel. vs reference)       */
#define B200_PREC_FP32 3 /* fp32-parity: each operand split in 3 bf16 planes, 6 MMAs per product    */
                         /* (tolerance 1e-5 rel. vs reference fp32)                                 */

#define B200_PLAN_TRAIN 0 /* forward(train/eval) + backward + optimizer */
#define B200_PLAN_INFER 1 /* eval-mode forward only (smaller workspace)  */

typedef struct b200_plan b200_plan;

typedef struct b200_plan_desc {
  int32_t batch;     /* rows per call (B >= 1; training needs B >= 2 like nn.BatchNorm1d) */
  int32_t precision; /* B200_PREC_* */
  int32_t kind;      /* B200_PLAN_* */
  int32_t reserved;
  float* params;     /* flat parameters [B200_PARAM_NUMEL] */
  float* grads;      /* flat gradients  [B200_PARAM_NUMEL]  (NULL for B200_PLAN_INFER) */
  float* exp_avg;    /* Adam first moment  (NULL if the optimizer step is never called) */
  float* exp_avg_sq; /* Adam second moment */
  float* running_mean[B200_NUM_BN];          /* BatchNorm1d buffers, [1024] each, in module order */
  float* running_var[B200_NUM_BN];
  int64_t* num_batches_tracked[B200_NUM_BN]; /* scalar each */
  void* workspace;   /* >= b200_workspace_bytes(batch, precision, kind), 1024-byte aligned */
  int64_t workspace_bytes;
} b200_plan_desc;

/* per-step optimizer hyper-parameters: torch.optim.Adam defaults of model/bilinear.py:60 plus the
 * LR schedule of util/config.py:21-22 evaluated by the caller (train_bilinear.py:66-70). */
typedef struct b200_adam_args {
  float lr;
  float beta1, beta2, eps;
  int64_t step;     /* Adam's per-parameter step AFTER this update (1 on the first update) */
  float max_norm;   /* clip_grad_norm_(max_norm) of train_bilinear.py:81; <= 0 disables clipping */
  float grad_scale; /* multiply gradients before clipping (1/world_size under data parallel) */
} b200_adam_args;

typedef struct b200_train_args {
  const float* x;        /* [B][32] */
  const float* y;        /* [B][48] */
  float* p_out;          /* [B][48] prediction of this step, or NULL */
  const uint8_t* masks;  /* [5][B][1024] dropout KEEP masks (1 = keep) shared with the oracle, or NULL */
  uint64_t seed;         /* Philox seed for in-kernel dropout when masks == NULL */
  uint64_t step;         /* global step counter (salts the Philox stream) */
  float momentum[B200_NUM_BN]; /* BatchNorm1d.momentum per layer; < 0 means None (cumulative average,
                                  as set by BilinearUnit.reset_statistics, model/bilinear.py:43-55) */
  float* loss_out;       /* scalar: mean squared error of this step (nn.MSELoss), or NULL */
  float* gnorm_out;      /* scalar: total gradient norm before clipping, or NULL */
} b200_train_args;

int b200_version(void);
const char* b200_last_error(void);

int64_t b200_param_numel(void);
int64_t b200_workspace_bytes(int32_t batch, int32_t precision, int32_t kind);

int b200_plan_create(const b200_plan_desc* desc, b200_plan** out);
void b200_plan_destroy(b200_plan* plan);

/* Re-derive the bf16 operand planes from the fp32 master parameters.  Call after anything other
 * than b200_train_step / b200_adam_step wrote `params` (init, load_state_dict, checkpoint resume,
 * model/bilinear.py:74-90). */
int b200_refresh_weights(b200_plan* plan, void* stream);

/* BilinearUnit.forward in eval mode (valid_bilinear.py:31,52): BN uses running statistics,
 * dropout is the identity. */
int b200_forward_eval(b200_plan* plan, const float* x, float* p, void* stream);

/* BilinearUnit.forward in training mode (train_bilinear.py:76): batch statistics, running-stat
 * update, dropout; keeps what backward needs inside the workspace.  Only x, p_out, masks, seed,
 * step and momentum of `args` are read. */
int b200_forward_train(b200_plan* plan, const b200_train_args* args, void* stream);

/* loss.backward() of train_bilinear.py:79 for an arbitrary upstream gradient dP [B][48]; overwrites
 * (does not accumulate into) the flat gradient buffer. */
int b200_backward(b200_plan* plan, const float* dP, void* stream);

/* One whole iteration of train_bilinear.py:75-83 -- zero_grad, forward, nn.MSELoss, backward,
 * clip_grad_norm_, Adam.step -- as ONE persistent cooperative kernel.
 * With update == 0 the optimizer phase is skipped (data parallel: all-reduce `grads`, then call
 * b200_adam_step). */
int b200_train_step(b200_plan* plan, const b200_train_args* targs, const b200_adam_args* aargs, int32_t update,
                    void* stream);

/* The same iteration fed from HOST memory, i.e. the body of the loop train_bilinear.py:62-88 including its
 * transfers: x_host [B][32] and y_host [B][48] (pinned memory for truly asynchronous copies) are copied to
 * staging buffers inside the workspace (train_bilinear.py:72-73), the step runs as one kernel, and the mean
 * loss and the pre-clip gradient norm are copied back to out_host[0], out_host[1] (train_bilinear.py:86-88).
 * targs->x / y / loss_out / gnorm_out are ignored.  Asynchronous: synchronise `stream` before reading
 * out_host. */
int b200_train_step_host(b200_plan* plan, const float* x_host, const float* y_host, const b200_train_args* targs,
                         const b200_adam_args* aargs, int32_t update, float* out_host, void* stream);

/* Eval-mode forward fed from / returning to HOST memory (valid_bilinear.py:39-52 with its transfers):
 * x_host [B][32] -> p_host [B][48].  Asynchronous like b200_train_step_host. */
int b200_forward_eval_host(b200_plan* plan, const float* x_host, float* p_host, void* stream);

/* ---- data parallel (new; the reference is single-process) -----------------
 * One process per GPU.  The ONE exchange step of an iteration -- the all-reduce of the flat gradient --
 * runs INSIDE the persistent kernel over NVLink / NVSwitch peer memory: every rank allocates its gradient
 * buffer with b200_comm_alloc (cudaMalloc + CUDA IPC handle), the 64-byte handles are exchanged by the host
 * (torch.distributed all_gather), peers are mapped with b200_comm_open, and b200_plan_set_comm hands the
 * table to a training plan whose `grads` IS this rank's buffer.  b200_train_step_dp then runs forward,
 * backward, reduce-scatter + all-gather over peer memory, clip_grad_norm_ and Adam (gradients x
 * aargs->grad_scale = 1/world) as one kernel; all ranks must call it for the same iteration.
 * BatchNorm uses each rank's own batch statistics (what DistributedDataParallel does). */
typedef struct b200_comm_desc {
  int32_t rank, world;      /* world <= 8 */
  uint32_t epoch;           /* 0, or the barrier epoch to resume from (shared by all plans of the process) */
  uint32_t reserved;
  void* buffers[8];         /* every rank's communication buffer (b200_comm_bytes() each), index = rank */
} b200_comm_desc;
int64_t b200_comm_bytes(void);   /* 4 * B200_PARAM_NUMEL + control block */
int b200_comm_alloc(int64_t bytes, void** dev_ptr, void* handle64);
int b200_comm_open(const void* handle64, void** peer_ptr);
int b200_comm_close(void* peer_ptr);
int b200_comm_free(void* dev_ptr);
int b200_plan_set_comm(b200_plan* plan, const b200_comm_desc* comm);
/* epoch_io (host, may be NULL): in = barrier epoch to use (0: the plan's own count), out = next epoch */
int b200_train_step_dp(b200_plan* plan, const b200_train_args* targs, const b200_adam_args* aargs, uint32_t* epoch_io,
                       void* stream);

/* clip_grad_norm_ + Adam.step on whatever is in `grads` (train_bilinear.py:81-83). */
int b200_adam_step(b200_plan* plan, const b200_adam_args* aargs, float* gnorm_out, void* stream);

/* ---- evaluation math ---------------------------------------------------------------------------
 * valid_bilinear.py:53-60: un-normalise prediction and ground truth with the train-set mean/std
 * ([3J] each, NULL = already in mm), per-pose sum over joints of the L2 distance, accumulated in
 * fp64 per action bucket (valid_bilinear.py:61-70).  pred, gt: [N][J][3] fp32.  J in {16, 17}.
 *   dist_out  [N] fp32 per-pose sums (may be NULL)
 *   action    [N] int32 bucket ids in [0, n_actions) (NULL: everything in bucket 0)
 *   acc_sum   [n_actions] fp64, acc_cnt [n_actions] int64: ACCUMULATED into (caller zeroes).
 * b200_pmpjpe additionally aligns each prediction to its ground truth with the optimal similarity
 * transform (rotation without reflection, uniform scale, translation -- Procrustes; not present in
 * the reference) before measuring. */
int b200_mpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n, int32_t joints,
               const int32_t* action, int32_t n_actions, float* dist_out, double* acc_sum, int64_t* acc_cnt,
               void* stream);
int b200_pmpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n, int32_t joints,
                const int32_t* action, int32_t n_actions, float* dist_out, double* acc_sum, int64_t* acc_cnt,
                void* stream);
/* both metrics in one pass over the data (384 B/pose read once). dist_out/pdist_out may be NULL. */
int b200_mpjpe_pmpjpe(const float* pred, const float* gt, const float* mean, const float* stddev, int64_t n,
                      int32_t joints, const int32_t* action, int32_t n_actions, float* dist_out, float* pdist_out,
                      double* acc_sum, double* pacc_sum, int64_t* acc_cnt, void* stream);

/* ---- introspection / tests --------------------------------------------------------------------
 * number of kernels this library has launched on the calling process so far */
int64_t b200_launch_count(void);
/* keep bits (dropout keep AND BatchNorm output > 0) of heavy layer `layer` from the last
 * training-mode forward of this plan: out [1024][ceil(B/32)] uint32, bit j of word w <-> row 32 w + j */
int b200_debug_keep_bits(b200_plan* plan, int32_t layer, uint32_t* out, void* stream);
/* per-phase durations (ns, from %globaltimer stamps) of the most recent launch on this plan; HOST
 * output arrays; synchronises the device.  kinds_out: 0 = GEMM phase, else 1 + elementwise kind;
 * tiles_out: GEMM tiles | epilogue kind << 24, or number of elementwise jobs. */
int b200_debug_phase_ns(b200_plan* plan, int64_t* ns_out, int32_t* kinds_out, int32_t* tiles_out, int32_t max_phases,
                        int32_t* count_out);
/* D[M][N] (fp32, ld = N) = A * B^T on the tcgen05 path.  Operands are bf16 plane tensors:
 *   a_mn == 0: A is [planes][M][K] (K contiguous);  a_mn == 1: A is [planes][K][M] (M contiguous)
 *   likewise for B with N.  bn = tile width (multiple of 16, <= 256; multiple of 64 if b_mn). */
int b200_debug_gemm(const uint16_t* a, const uint16_t* b, float* d, int32_t m, int32_t n, int32_t k, int32_t planes,
                    int32_t a_mn, int32_t b_mn, int32_t bn, void* scratch, int64_t scratch_bytes, void* stream);
/* fp32 [rows][cols] -> bf16 planes [planes][rows][cols] */
int b200_debug_split(const float* src, uint16_t* dst, int64_t numel, int32_t planes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BILINEAR_B200_H_ */
