// CTA-pair build of the persistent kernel (tcgen05 cta_group::2, M = 256; the default for large-batch plans).
// Own translation unit: ptxas 12.9 crashes when two instantiations of the kernel share a module.
// -DB200_PAIR_E8=1 builds it with EIGHT epilogue warps in role groups with their own register budgets (mega.cu,
// B200_MEGA_E8) -- measured 3.08-3.11 ms per batch-65536 step against 2.90-2.93 ms for this build.
#define B200_MEGA_CG 2
#ifdef B200_PAIR_E8
#define B200_MEGA_E8 1
#endif
#include "mega.cu"
