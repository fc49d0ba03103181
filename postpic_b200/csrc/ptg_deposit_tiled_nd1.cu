// The 1-D kernel instances of ptg_deposit_tiled.cu as a translation unit of their own (build time: see the note at that
// file's dispatcher).
#define PTG_ND 1
#include "ptg_deposit_tiled.cu"
