// Tile descriptors of the register-blocked Qk stage kernel (row_kernel.cuh), built once on the
// host from the generic unique-face list of partition.h.
//
// In that kernel a tile cell is worked on by N1 = k+1 threads; thread l of a cell holds Gauss row
// l (then Gauss column l) of the cell in registers.  Every face is solved along +e_x / +e_y by
// exactly one agent:
//   * the HIGH face (right, top) of a tile cell: by the cell's own thread of that row / column.
//     nbhi[slot][dir] says where the neighbour's trace comes from: another tile cell (whose low
//     face then receives the flux), a ghost-trace slot (neighbour outside the tile, or a periodic
//     partner), or a physical boundary;
//   * a LOW face (left, bottom) whose neighbour is not a tile cell reached through an ordinary
//     interior face: by an "L job" of the block's two extra warps;
//   * ghost traces (the low-face trace of the cell beyond a high tile-edge face) are produced by
//     "G jobs" of the extra warps from the staged halo cells.
// Which cell is the reference's "plus" side (the one MeshWorker::loop integrates the face from,
// reference src/assemble_explicit.cc:440-451, or both for periodic pairs, src_mpi 186-260) is
// carried per face so the Riemann problem is posed exactly as in the reference.
#pragma once

#include <vector>

namespace dflo
{
   // tile shape (cells) per N1; TC = tx*ty is a multiple of 32 so that TC*N1 threads are whole warps
   constexpr int row_tx (int n1) { return n1 == 2 ? 8 : 8; }
#ifndef DFLO_ROW_TY
#define DFLO_ROW_TY 4
#endif
   constexpr int row_ty (int n1) { return n1 == 2 ? 8 : DFLO_ROW_TY; }
   constexpr int row_tc (int n1) { return row_tx (n1) * row_ty (n1); }
   constexpr int row_nh (int n1) { return 2 * (row_tx (n1) + row_ty (n1)); } // staged halo cells = L-job = G-job capacity

   // descriptor of one tile, in ints:
   //   [0] c0  [1] ncb  [2] nh  [3] nL  [4] nG  [5..7] 0
   //   halo[nh_cap]            local cell ids staged after the tile cells (su slot TC + i)
   //   nbhi[tc][2]             code >= 0: (index & 0xffff) | plus_own << 16, index < TC: tile slot, else TC + ghost slot
   //                           code <  0: -1 - local boundary face
   //   ljob[nl_cap][2]         { (slot*2+dir) | plus_own << 16 | flip << 17,  nb: su slot >= 0 or -1 - local boundary face }
   //   gjob[ng_cap]            su slot | dir << 16 | flip << 17
   enum { ROWD_HDR = 8, ROWD_PLUS = 1 << 16, ROWD_FLIP = 1 << 17 };
   constexpr int rowd_off_halo () { return ROWD_HDR; }
   constexpr int rowd_off_nbhi (int nh) { return ROWD_HDR + nh; }
   constexpr int rowd_off_ljob (int tc, int nh) { return ROWD_HDR + nh + 2 * tc; }
   constexpr int rowd_off_gjob (int tc, int nh) { return ROWD_HDR + nh + 2 * tc + 2 * nh; }
   constexpr int rowd_ints (int tc, int nh) { return (ROWD_HDR + nh + 2 * tc + 2 * nh + nh + 3) / 4 * 4; } // 16-byte multiple
}
