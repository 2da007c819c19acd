// halves of 8 warps (512-thread CTA, <= 128 registers), 4 matrix rows per pipeline stage
#define DUO_HWN 4
#define DUO_BK 4
#define DUO_ENTRY launch_prune_duo_h4k4
#include "tu_prune_duo.inc"
