// two 256-thread CTAs per SM (<= 128 registers), BN = 32*TNW with TNW <= 2, BK = 4
#define RS_WN 4
#define RS_BK 4
#define RS_MINB 2
#define RS_MAXTNW 2
#define RS_ENTRY launch_prune_resident_wn4x2
#include "tu_prune_resident.inc"
