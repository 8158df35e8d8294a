// K1 TMA-staged kernel with 16 frames per tile (one 16-warp CTA per SM): the same source as stft_tma.cu compiled with
// DAM_TMA_FT = 16 and exported as dam_k1_launch_tma16.  Every store instruction then writes two 64-byte row segments
// instead of four 32-byte ones: measured +30 % on output rows that are not a multiple of 8 frames (T = 173 dense:
// 0.867 -> 0.666 ms) and +3.6 % on long sector-aligned rows at hop 1024 (T = 216: 1.127 -> 1.089 ms), -4 % on the
// hop-512 bench shape and on short rows (more of the last tile is wasted) -- the dispatcher in stft_db.cu picks.
#define DAM_TMA_FT 16
#include "stft_tma.cu"
