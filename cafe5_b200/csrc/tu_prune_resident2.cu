// two 128-thread CTAs per SM, BN = 16*TNW, BK = 4
#define RS_WN 2
#define RS_BK 4
#define RS_MINB 2
#define RS_ENTRY launch_prune_resident_wn2
#define RS_JOBS_ENTRY launch_prune_resident_wn2_tables
#include "tu_prune_resident.inc"
