// CTA-pair (tcgen05 cta_group::2, clusters of two) instantiation of the persistent kernel.
// Same source as mega.cu; a separate module because ptxas 12.9 crashes with both variants in one.
#define B200_MEGA_CG 2
#include "mega.cu"
