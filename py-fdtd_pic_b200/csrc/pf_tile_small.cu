// pf_tile_small.cu -- the tile engine once more, as namespace pf::small with 256-cell tiles (128 threads, k <= 64): the latency
// variant pf_run_pass uses for a single grid too short to fill the GPU with 1024-cell tiles (see the note at the top of
// pf_tile.cu).  Same source, same arithmetic, same results: tiling never changes a bit (tests/test_gpu_parity.py).
#define PF_TILE_SECONDARY
#define PF_TILE_CELLS 256
#define PF_TILE_KMAX 64
#define PF_TILE_C_FREE 2
#define PF_TILE_MINBLOCKS_FREE 2
#define PF_TILE_MINBLOCKS_F32 2
#include "pf_tile.cu"
