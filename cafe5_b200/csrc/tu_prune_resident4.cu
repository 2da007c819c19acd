// one 256-thread CTA per SM, BN = 32*TNW, BK = 8
#define RS_WN 4
#define RS_BK 8
#define RS_MINB 1
#define RS_ENTRY launch_prune_resident_wn4
#include "tu_prune_resident.inc"
