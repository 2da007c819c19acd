// the default geometry (two 128-thread CTAs per SM, BK = 4) with clock stamps around the phases of the chunk loop (VARIANT 2)
#define RS_WN 2
#define RS_BK 4
#define RS_MINB 2
#define RS_VARIANT 2
#define RS_ENTRY launch_prune_resident_wn2probe
#include "tu_prune_resident.inc"
