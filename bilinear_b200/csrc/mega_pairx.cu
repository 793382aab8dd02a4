// CTA-pair build of the persistent kernel WITH helper warps: 24 warps per CTA, the elementwise jobs of a phase
// (large-batch BatchNorm apply passes) run on warps 8-23 concurrently with the GEMM tiles that consume their
// output, chunk by chunk (see mega.cu, B200_MEGA_HELPERS).  Own translation unit: see mega_pair.cu.
#define B200_MEGA_CG 2
#define B200_MEGA_HELPERS 1
#include "mega.cu"
