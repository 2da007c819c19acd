// halves of 4 warps (256-thread CTA), 8 matrix rows per pipeline stage
#define DUO_HWN 2
#define DUO_BK 8
#define DUO_ENTRY launch_prune_duo_h2
#include "tu_prune_duo.inc"
