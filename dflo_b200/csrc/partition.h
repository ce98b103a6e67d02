// Cell sharding of the flat mesh across ranks (one process per GPU).  Cells are split by global
// cell id into `world` contiguous ranges (what dflo's src_mpi tree gets from p4est, reference
// src_mpi/claw.cc:331-340, reduced to the structured case).  Each rank holds its owned cells plus
// `layers` layers of ghost cells in the face-neighbour graph:
//   layers = 1: the reference's one-layer halo, enough when no TVB limiter is active;
//   layers = 2: lets the stage kernel redundantly update the first ghost layer so that the TVB
//               limiter of owned cells sees its neighbours' post-update means without a second
//               exchange: still ONE halo exchange per RK stage.
// Local cell order: owned (tile-major, see below) | ghost layer 1 grouped by owner rank | ghost
// layer 2 grouped by owner rank, each ghost group in global order, so every peer's data lands in
// at most two contiguous ranges and needs no unpack kernel.
//
// Tiles.  The stage kernel works on one TILE of cells per thread block: a run of consecutive local
// cells whose DoFs are one contiguous block of memory (one bulk async copy into shared memory).
// Owned cells are therefore renumbered tile-major: on a uniform Cartesian lattice (every mesh the
// reference's TVB/Pk paths accept, reference src/parameters.cc:536-550) tiles are tx x ty patches
// of the lattice; otherwise they are grown greedily through the face-neighbour graph.  For every
// tile the host flattens, once, (a) the list of distinct cells outside the tile that touch it
// (its halo, staged next to the tile in shared memory) and (b) the list of UNIQUE faces: an
// interior face whose two cells sit in the same tile appears once, from the cell that
// MeshWorker::loop visits it from (reference src/assemble_explicit.cc:440-451), so its Riemann
// problem is solved once per tile instead of once per adjacent cell.  Correctness never depends
// on how cells were grouped, only the amount of sharing does.
#pragma once

#include "../../include/dflo_b200.h"
#include "row_desc.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace dflo
{
   struct HaloPeer
   {
      int rank;
      std::vector<int> send_cells[2]; // local (owned) cell ids to send, per ghost layer of the peer
      int recv_start[2], recv_count[2]; // contiguous local ranges receiving from this peer
      int dst_start[2];                 // where send_cells[k] land in the PEER's local cell numbering (its recv_start for us)
   };

   struct LocalMesh
   {
      int rank, world, layers;
      int n_global;
      int begin, end;            // owned global range
      int n_owned, n_ghost1, n_ghost2, n_local;
      int n_compute;             // cells the stage kernel updates: owned (+ ghost layer 1 when layers == 2)
      std::vector<int> l2g;      // local -> global cell
      std::vector<int> nbr;      // [n_local][4] local cell / -1-local bface; self for unknown neighbours
      std::vector<uint8_t> fflags;
      std::vector<double> geom;  // [n_local][4] x0 y0 hx hy
      std::vector<double> verts; // [n_local][8] cell vertices (mapping = q1; empty otherwise)
      std::vector<unsigned char> nbr_face; // [n_local][4] the neighbour's local number of the shared face (mapping = q1)
      // faces with a hanging node: hang_of[cell*4+f] = index or -1; per face
      // { fine cell 0, its face, runs backwards?, fine cell 1, its face, runs backwards? } in local cell numbers
      std::vector<int> hang_of, hang;
      std::vector<int> bf_global, bf_id; // local boundary faces (of the computed cells) -> global bface, boundary id
      std::vector<int> bf_cell, bf_face; // local cell, face
      std::vector<HaloPeer> peers;

      // tiles of the computed cells [0, n_compute): owned tiles first, then ghost-layer-1 tiles
      int tile_cells = 0, tile_halo_max = 0;
      int n_tiles = 0, n_tiles_owned = 0;
      std::vector<int> tile_start;   // [n_tiles+1] first local cell of each tile
      std::vector<int> halo_start;   // [n_tiles+1] into halo_cells
      std::vector<int> halo_cells;   // local cell ids staged after the tile's own cells
      std::vector<int> job_start;    // [n_tiles+1] into jobs (in units of jobs)
      std::vector<int> jobs;         // 4 ints per unique face, see FaceJob in kernels.cuh

      // descriptors of the register-blocked Qk stage kernel (row_desc.h), rowdesc_stride ints per tile
      std::vector<int> rowdesc;
      int rowdesc_stride = 0;
      // fused halo exchange (p2p_halo.cuh): per tile, the owned cells a peer needs
      std::vector<int> send_entries; // 3 ints per entry: local cell, peer index (into peers), destination cell on the peer
      int n_send_tiles = 0;
   };

   constexpr int JOB_SHARED_FLAG = 8; // job flag beside DFLO_FACE_* (kernels.cuh JOB_SHARED): the flux also serves the neighbour slot

   inline void partition_range (int n, int world, int r, int &b, int &e)
   {
      const int64_t q = n / world, rem = n % world;
      b = (int) (r * q + std::min<int64_t> (r, rem));
      e = b + (int) q + (r < rem ? 1 : 0);
   }

   inline int owner_of (int n, int world, int cell)
   {
      const int64_t q = n / world, rem = n % world;
      const int64_t split = rem * (q + 1);
      if (cell < split) return (int) (cell / (q + 1));
      return (int) (rem + (cell - split) / (q > 0 ? q : 1));
   }

   // ghost layers of rank r: g1 = non-owned face neighbours of owned cells, g2 = further
   // non-owned face neighbours of g1 (sorted by (owner, global id))
   inline void ghost_layers (const dflo_flat_mesh &m, int world, int r, int layers, std::vector<int> &g1, std::vector<int> &g2)
   {
      int b, e;
      partition_range (m.n_cells, world, r, b, e);
      g1.clear ();
      g2.clear ();
      std::vector<char> mark (m.n_cells, 0);
      // all face neighbours of a cell: `neighbor`, plus the second fine cell behind a face with a hanging node
      std::vector<int> second (m.n_hanging_faces > 0 ? 4 * (size_t) m.n_cells : 0, -1);
      for (int h = 0; h < m.n_hanging_faces; ++h) second[4 * (size_t) m.hanging[6 * h] + m.hanging[6 * h + 1]] = m.hanging[6 * h + 4];
      auto visit = [&] (int c, char layer, std::vector<int> &out) {
         for (int f = 0; f < 4; ++f)
            for (int k = 0; k < 2; ++k)
            {
               const int nb = k == 0 ? m.neighbor[4 * (size_t) c + f] : (second.empty () ? -1 : second[4 * (size_t) c + f]);
               if (nb >= 0 && (nb < b || nb >= e) && !mark[nb])
               {
                  mark[nb] = layer;
                  out.push_back (nb);
               }
            }
      };
      for (int c = b; c < e; ++c) visit (c, 1, g1);
      if (layers >= 2)
      {
         const std::vector<int> first (g1);
         for (int c : first) visit (c, 2, g2);
      }
      std::sort (g1.begin (), g1.end ()); // contiguous ownership => sorted by (owner, id)
      std::sort (g2.begin (), g2.end ());
   }

   // Tile-major order of the owned cells [b, e): returns the global cell ids in their new local
   // order and the tile boundaries (prefix offsets into that order).
   inline void order_owned_cells (const dflo_flat_mesh &m, int b, int e, int tx, int ty, std::vector<int> &order,
                                  std::vector<int> &tile_start)
   {
      const int n = e - b, tc = tx * ty;
      order.clear ();
      tile_start.assign (1, 0);
      // uniform lattice?
      const double hx = m.cell_size[0], hy = m.cell_size[1];
      double xmin = m.cell_origin[0], ymin = m.cell_origin[1];
      bool uniform = hx > 0 && hy > 0;
      for (int c = 0; c < m.n_cells && uniform; ++c)
      {
         uniform = std::fabs (m.cell_size[2 * (size_t) c] - hx) <= 1e-9 * hx && std::fabs (m.cell_size[2 * (size_t) c + 1] - hy) <= 1e-9 * hy;
         xmin = std::min (xmin, m.cell_origin[2 * (size_t) c]);
         ymin = std::min (ymin, m.cell_origin[2 * (size_t) c + 1]);
      }
      if (uniform && tc > 1)
      {
         struct Key { int64_t tile; int iy, ix, cell; };
         std::vector<Key> keys (n);
         for (int i = 0; i < n; ++i)
         {
            const int c = b + i;
            const int ix = (int) std::llround ((m.cell_origin[2 * (size_t) c] - xmin) / hx);
            const int iy = (int) std::llround ((m.cell_origin[2 * (size_t) c + 1] - ymin) / hy);
            keys[i] = Key{((int64_t) (iy / ty) << 32) | (int64_t) (ix / tx), iy, ix, c};
         }
         std::sort (keys.begin (), keys.end (), [] (const Key &p, const Key &q) {
            if (p.tile != q.tile) return p.tile < q.tile;
            if (p.iy != q.iy) return p.iy < q.iy;
            if (p.ix != q.ix) return p.ix < q.ix;
            return p.cell < q.cell;
         });
         for (int i = 0; i < n; ++i)
         {
            // a new tile starts at every new lattice patch, and whenever a patch overflows (two
            // cells on one lattice site: overlapping blocks of a malformed mesh)
            if (i > 0 && (keys[i].tile != keys[i - 1].tile || i - tile_start.back () >= tc)) tile_start.push_back (i);
            order.push_back (keys[i].cell);
         }
         tile_start.push_back (n);
         return;
      }
      // general meshes: grow tiles breadth-first through the face-neighbour graph
      std::vector<char> taken (n, 0);
      std::vector<int> queue;
      for (int seed = 0; seed < n; ++seed)
      {
         if (taken[seed]) continue;
         queue.assign (1, seed);
         taken[seed] = 1;
         const int first = (int) order.size ();
         for (size_t head = 0; head < queue.size (); ++head)
         {
            const int c = b + queue[head];
            order.push_back (c);
            for (int f = 0; f < 4 && (int) queue.size () < tc; ++f)
            {
               const int nb = m.neighbor[4 * (size_t) c + f];
               if (nb >= b && nb < e && !taken[nb - b])
               {
                  taken[nb - b] = 1;
                  queue.push_back (nb - b);
               }
            }
         }
         tile_start.push_back (first + (int) queue.size ());
      }
   }

   // Sharded contexts: tiles that touch the partition cut (they own cells a peer needs and read
   // ghost cells) come first in the tile order, so their results travel to the peers while the
   // interior tiles are still being worked on.
   // first = false puts them LAST instead: when the exchange is a kernel of its own whose wait is left to the tiles
   // that read ghost cells, those tiles should come up when the peers' data has long arrived.
   inline void boundary_tiles_first (const dflo_flat_mesh &m, int b, int e, std::vector<int> &order, std::vector<int> &tile_start, bool first = true)
   {
      const int nt = (int) tile_start.size () - 1;
      std::vector<char> cut (nt, 0);
      for (int t = 0; t < nt; ++t)
         for (int i = tile_start[t]; i < tile_start[t + 1] && !cut[t]; ++i)
            for (int f = 0; f < 4; ++f)
            {
               const int nb = m.neighbor[4 * (size_t) order[i] + f];
               if (nb >= 0 && (nb < b || nb >= e)) cut[t] = 1;
            }
      std::vector<int> o2, ts (1, 0);
      for (int k = 0; k < 2; ++k)
         for (int t = 0; t < nt; ++t)
            if (cut[t] == (first ? 1 - k : k))
            {
               o2.insert (o2.end (), order.begin () + tile_start[t], order.begin () + tile_start[t + 1]);
               ts.push_back ((int) o2.size ());
            }
      order.swap (o2);
      tile_start.swap (ts);
   }

   // Unique-face job lists and tile halos (needs L.nbr / L.fflags / L.tile_start filled in)
   inline void build_tile_jobs (LocalMesh &L)
   {
      L.n_tiles = (int) L.tile_start.size () - 1;
      L.halo_start.assign (1, 0);
      L.job_start.assign (1, 0);
      L.halo_cells.clear ();
      L.jobs.clear ();
      std::vector<int> halo_slot (L.n_local, -1);
      for (int t = 0; t < L.n_tiles; ++t)
      {
         const int c0 = L.tile_start[t], c1 = L.tile_start[t + 1];
         const size_t h0 = L.halo_cells.size ();
         for (int cell = c0; cell < c1; ++cell)
            for (int f = 0; f < 4; ++f)
            {
               const int nb = L.nbr[4 * (size_t) cell + f];
               const int fl = L.fflags[4 * (size_t) cell + f];
               int slot_b = -1, flags = fl;
               if (nb >= 0)
               {
                  const bool inside = nb >= c0 && nb < c1;
                  if (inside && !(fl & DFLO_FACE_PERIODIC))
                  {
                     if (!(fl & DFLO_FACE_OWNER)) continue; // the owner's job covers this side too
                     slot_b = nb - c0;
                     flags |= JOB_SHARED_FLAG;
                  }
                  else if (inside)
                     slot_b = nb - c0; // periodic partner in the same tile: both sides integrate
                  else
                  {
                     if (halo_slot[nb] < 0 && (int) (L.halo_cells.size () - h0) < L.tile_halo_max)
                     {
                        halo_slot[nb] = (int) (L.halo_cells.size () - h0);
                        L.halo_cells.push_back (nb);
                     }
                     // beyond the staged halo capacity the kernel gathers from global memory
                     slot_b = halo_slot[nb] >= 0 ? L.tile_cells + halo_slot[nb] : -1;
                  }
               }
               L.jobs.push_back ((cell - c0) * 4 + f);
               L.jobs.push_back (nb);
               L.jobs.push_back (slot_b);
               L.jobs.push_back (flags);
            }
         for (size_t h = h0; h < L.halo_cells.size (); ++h) halo_slot[L.halo_cells[h]] = -1;
         L.halo_start.push_back ((int) L.halo_cells.size ());
         L.job_start.push_back ((int) (L.jobs.size () / 4));
      }
   }

   // Does tile t respect the capacities of the row kernel (staged halo cells, L jobs, G jobs <= cap)?
   inline bool row_tile_fits (const LocalMesh &L, int t, int cap)
   {
      int nL = 0, nG = 0;
      for (int j = L.job_start[t]; j < L.job_start[t + 1]; ++j)
      {
         const int a = L.jobs[4 * (size_t) j], nb = L.jobs[4 * (size_t) j + 1], slot_b = L.jobs[4 * (size_t) j + 2];
         const int flags = L.jobs[4 * (size_t) j + 3];
         if (flags & JOB_SHARED_FLAG) continue;
         const bool high = (a & 1) != 0;
         if (nb >= 0)
         {
            if (slot_b < 0) return false; // neighbour beyond the staged halo
            if (high) ++nG; else ++nL;
         }
         else if (!high)
            ++nL;
      }
      return nL <= cap && nG <= cap;
   }

   // Tile jobs for the row kernel: tiles that overflow its capacities (ragged tiles of general
   // meshes, the strip-shaped tiles of a redundantly updated ghost layer) are halved until they fit.
   inline void build_tiles_for_row (LocalMesh &L, int cap)
   {
      for (;;)
      {
         build_tile_jobs (L);
         std::vector<int> ts (1, 0);
         int owned = 0;
         bool split = false;
         for (int t = 0; t < L.n_tiles; ++t)
         {
            const int c0 = L.tile_start[t], c1 = L.tile_start[t + 1];
            int pieces = 1;
            if (c1 - c0 > 1 && !row_tile_fits (L, t, cap))
            {
               ts.push_back (c0 + (c1 - c0) / 2);
               pieces = 2;
               split = true;
            }
            ts.push_back (c1);
            if (t < L.n_tiles_owned) owned += pieces;
         }
         if (!split) return;
         L.tile_start = ts;
         L.n_tiles_owned = owned;
      }
   }

   // Row-kernel descriptors (row_desc.h) from the generic unique-face lists
   inline bool build_row_desc (LocalMesh &L, int tc, int nh, std::string &err)
   {
      const int stride = rowd_ints (tc, nh);
      L.rowdesc_stride = stride;
      L.rowdesc.assign ((size_t) std::max (1, L.n_tiles) * stride, 0);
      const int UNSET = 0x7fffffff;
      for (int t = 0; t < L.n_tiles; ++t)
      {
         int *d = &L.rowdesc[(size_t) t * stride];
         const int c0 = L.tile_start[t], ncb = L.tile_start[t + 1] - c0;
         const int nhl = L.halo_start[t + 1] - L.halo_start[t];
         int *halo = d + rowd_off_halo (), *nbhi = d + rowd_off_nbhi (nh), *lj = d + rowd_off_ljob (tc, nh), *gj = d + rowd_off_gjob (tc, nh);
         for (int i = 0; i < nhl; ++i) halo[i] = L.halo_cells[L.halo_start[t] + i];
         for (int i = 0; i < 2 * tc; ++i) nbhi[i] = UNSET;
         int nL = 0, nG = 0;
         for (int j = L.job_start[t]; j < L.job_start[t + 1]; ++j)
         {
            const int a = L.jobs[4 * (size_t) j], nb = L.jobs[4 * (size_t) j + 1], slot_b = L.jobs[4 * (size_t) j + 2];
            const int flags = L.jobs[4 * (size_t) j + 3];
            const int sa = a >> 2, f = a & 3, dir = f >> 1;
            const bool high = (f & 1) != 0;
            const bool plus_own = nb < 0 || (flags & (DFLO_FACE_OWNER | DFLO_FACE_PERIODIC));
            const int flip = (flags & DFLO_FACE_FLIP) ? ROWD_FLIP : 0;
            if (flags & JOB_SHARED_FLAG)
            {
               // listed from the owner (plus) cell sa; the low-side cell's thread solves it
               const int lo = high ? sa : slot_b, hi = high ? slot_b : sa;
               nbhi[2 * lo + dir] = hi | (high ? ROWD_PLUS : 0);
            }
            else if (nb >= 0)
            {
               if (slot_b < 0 || nL >= nh || nG >= nh)
               {
                  err = "internal: row-kernel tile exceeds its capacities";
                  return false;
               }
               if (high)
               {
                  gj[nG] = slot_b | (dir << 16) | flip;
                  nbhi[2 * sa + dir] = (tc + nG) | (plus_own ? ROWD_PLUS : 0);
                  ++nG;
               }
               else
               {
                  lj[2 * nL] = (2 * sa + dir) | (plus_own ? ROWD_PLUS : 0) | flip;
                  lj[2 * nL + 1] = slot_b;
                  ++nL;
               }
            }
            else if (high)
               nbhi[2 * sa + dir] = nb; // -1 - local boundary face
            else
            {
               if (nL >= nh)
               {
                  err = "internal: row-kernel tile exceeds its capacities";
                  return false;
               }
               lj[2 * nL] = (2 * sa + dir) | ROWD_PLUS;
               lj[2 * nL + 1] = nb;
               ++nL;
            }
         }
         for (int i = 0; i < 2 * ncb; ++i)
            if (nbhi[i] == UNSET)
            {
               err = "internal: a high face of a tile cell has no agent";
               return false;
            }
         d[0] = c0;
         d[1] = ncb;
         d[2] = nhl;
         d[3] = nL;
         d[4] = nG;
      }
      return true;
   }

   inline bool build_local_mesh (const dflo_flat_mesh &m, int rank, int world, int layers, int tile_x, int tile_y,
                                 LocalMesh &L, std::string &err, bool row = false, bool boundary_first = true)
   {
      if (world < 1 || rank < 0 || rank >= world || m.n_cells < world)
      {
         err = "invalid rank/world for this mesh";
         return false;
      }
      L = LocalMesh ();
      L.rank = rank;
      L.world = world;
      L.layers = world == 1 ? 0 : layers;
      L.n_global = m.n_cells;
      partition_range (m.n_cells, world, rank, L.begin, L.end);
      L.n_owned = L.end - L.begin;
      std::vector<int> g1, g2;
      if (world > 1) ghost_layers (m, world, rank, layers, g1, g2);
      L.n_ghost1 = g1.size ();
      L.n_ghost2 = g2.size ();
      L.n_local = L.n_owned + L.n_ghost1 + L.n_ghost2;
      L.n_compute = L.n_owned + (L.layers >= 2 ? L.n_ghost1 : 0);
      L.l2g.resize (L.n_local);
      L.tile_cells = tile_x * tile_y;
      L.tile_halo_max = 2 * (tile_x + tile_y);
      std::vector<int> order;
      order_owned_cells (m, L.begin, L.end, tile_x, tile_y, order, L.tile_start);
      if (world > 1) boundary_tiles_first (m, L.begin, L.end, order, L.tile_start, boundary_first);
      L.n_tiles_owned = (int) L.tile_start.size () - 1;
      std::vector<int> owned_g2l (L.n_owned);
      for (int i = 0; i < L.n_owned; ++i)
      {
         L.l2g[i] = order[i];
         owned_g2l[order[i] - L.begin] = i;
      }
      // ghost-layer-1 cells that are updated redundantly: tiles of consecutive cells
      for (int c = L.n_owned; c < L.n_compute; c += L.tile_cells) L.tile_start.push_back (std::min (c + L.tile_cells, L.n_compute));
      for (int i = 0; i < L.n_ghost1; ++i) L.l2g[L.n_owned + i] = g1[i];
      for (int i = 0; i < L.n_ghost2; ++i) L.l2g[L.n_owned + L.n_ghost1 + i] = g2[i];
      // global -> local lookup for the cells we hold
      auto g2l = [&] (int g) -> int {
         if (g >= L.begin && g < L.end) return owned_g2l[g - L.begin];
         auto it = std::lower_bound (g1.begin (), g1.end (), g);
         if (it != g1.end () && *it == g) return L.n_owned + (int) (it - g1.begin ());
         it = std::lower_bound (g2.begin (), g2.end (), g);
         if (it != g2.end () && *it == g) return L.n_owned + L.n_ghost1 + (int) (it - g2.begin ());
         return -1;
      };
      L.nbr.resize (4 * (size_t) L.n_local);
      L.fflags.resize (4 * (size_t) L.n_local);
      L.geom.resize (4 * (size_t) L.n_local);
      if (m.cell_vertices && m.neighbor_face)
      {
         L.verts.resize (8 * (size_t) L.n_local);
         L.nbr_face.resize (4 * (size_t) L.n_local);
      }
      for (int l = 0; l < L.n_local; ++l)
      {
         const int g = L.l2g[l];
         if (!L.verts.empty ())
         {
            for (int i = 0; i < 8; ++i) L.verts[8 * (size_t) l + i] = m.cell_vertices[8 * (size_t) g + i];
            for (int f = 0; f < 4; ++f) L.nbr_face[4 * (size_t) l + f] = m.neighbor_face[4 * (size_t) g + f];
         }
         L.geom[4 * (size_t) l + 0] = m.cell_origin[2 * (size_t) g];
         L.geom[4 * (size_t) l + 1] = m.cell_origin[2 * (size_t) g + 1];
         L.geom[4 * (size_t) l + 2] = m.cell_size[2 * (size_t) g];
         L.geom[4 * (size_t) l + 3] = m.cell_size[2 * (size_t) g + 1];
         for (int f = 0; f < 4; ++f)
         {
            const int nb = m.neighbor[4 * (size_t) g + f];
            L.fflags[4 * (size_t) l + f] = m.face_flags[4 * (size_t) g + f];
            if (nb >= 0)
            {
               const int ln = g2l (nb);
               if (ln < 0 && l < L.n_compute)
               {
                  err = "internal: computed cell with a neighbour outside the halo";
                  return false;
               }
               L.nbr[4 * (size_t) l + f] = ln >= 0 ? ln : l;
            }
            else if (l < L.n_compute)
            {
               L.nbr[4 * (size_t) l + f] = -1 - (int) L.bf_global.size ();
               L.bf_global.push_back (-1 - nb);
               L.bf_id.push_back (m.bface_id[-1 - nb]);
               L.bf_cell.push_back (l);
               L.bf_face.push_back (f);
            }
            else
               L.nbr[4 * (size_t) l + f] = l; // never evaluated
         }
      }
      if (m.n_hanging_faces > 0)
      {
         std::vector<int> g2l_all (m.n_cells, -1);
         for (int l = 0; l < L.n_local; ++l) g2l_all[L.l2g[l]] = l;
         L.hang_of.assign (4 * (size_t) L.n_local, -1);
         for (int h = 0; h < m.n_hanging_faces; ++h)
         {
            const int *e = m.hanging + 6 * (size_t) h;
            const int lc = g2l_all[e[0]];
            if (lc < 0 || lc >= L.n_compute) continue; // the coarse cell is not updated on this rank
            L.hang_of[4 * (size_t) lc + e[1]] = (int) (L.hang.size () / 6);
            for (int k = 0; k < 2; ++k)
            {
               const int lf = g2l_all[e[2 + 2 * k]];
               if (lf < 0)
               {
                  err = "internal: a fine cell behind a hanging node is outside the halo";
                  return false;
               }
               L.hang.push_back (lf);
               L.hang.push_back (e[3 + 2 * k]);
               L.hang.push_back ((m.face_flags[4 * (size_t) e[2 + 2 * k] + e[3 + 2 * k]] & DFLO_FACE_FLIP) ? 1 : 0);
            }
         }
         if (L.hang.empty ()) L.hang.assign (6, 0);
      }
      if (row)
      {
         build_tiles_for_row (L, L.tile_halo_max);
         if (!build_row_desc (L, L.tile_cells, L.tile_halo_max, err)) return false;
      }
      else
         build_tile_jobs (L);
      if (world == 1) return true;

      // Halo lists.  What we receive: our ghosts, grouped by owner.  What we send to peer p: the
      // cells of p's ghost layers that we own, in p's local order (global order).
      std::vector<HaloPeer> peers (world);
      for (int p = 0; p < world; ++p)
      {
         peers[p].rank = p;
         for (int k = 0; k < 2; ++k) peers[p].recv_start[k] = peers[p].recv_count[k] = peers[p].dst_start[k] = 0;
      }
      for (int k = 0; k < 2; ++k)
      {
         const std::vector<int> &g = k == 0 ? g1 : g2;
         const int base = L.n_owned + (k == 0 ? 0 : L.n_ghost1);
         for (size_t i = 0; i < g.size (); ++i)
         {
            const int p = owner_of (m.n_cells, world, g[i]);
            if (peers[p].recv_count[k] == 0) peers[p].recv_start[k] = base + (int) i;
            peers[p].recv_count[k]++;
         }
      }
      for (int p = 0; p < world; ++p)
      {
         if (p == rank) continue;
         std::vector<int> pg1, pg2;
         ghost_layers (m, world, p, layers, pg1, pg2);
         int pb, pe;
         partition_range (m.n_cells, world, p, pb, pe);
         for (size_t i = 0; i < pg1.size (); ++i)
            if (pg1[i] >= L.begin && pg1[i] < L.end)
            {
               if (peers[p].send_cells[0].empty ()) peers[p].dst_start[0] = (pe - pb) + (int) i;
               peers[p].send_cells[0].push_back (owned_g2l[pg1[i] - L.begin]);
            }
         for (size_t i = 0; i < pg2.size (); ++i)
            if (pg2[i] >= L.begin && pg2[i] < L.end)
            {
               if (peers[p].send_cells[1].empty ()) peers[p].dst_start[1] = (pe - pb) + (int) pg1.size () + (int) i;
               peers[p].send_cells[1].push_back (owned_g2l[pg2[i] - L.begin]);
            }
      }
      for (int p = 0; p < world; ++p)
         if (p != rank
             && (peers[p].recv_count[0] || peers[p].recv_count[1] || !peers[p].send_cells[0].empty () || !peers[p].send_cells[1].empty ()))
            L.peers.push_back (peers[p]);
      if (row)
      {
         // per owned tile: which of its cells go where (fused exchange), and whether it reads ghosts
         std::vector<std::vector<int>> per_tile (L.n_tiles);
         std::vector<int> tile_of (L.n_owned, 0);
         for (int t = 0; t < L.n_tiles_owned; ++t)
            for (int c = L.tile_start[t]; c < L.tile_start[t + 1]; ++c) tile_of[c] = t;
         for (size_t pi = 0; pi < L.peers.size (); ++pi)
            for (int k = 0; k < 2; ++k)
               for (size_t j = 0; j < L.peers[pi].send_cells[k].size (); ++j)
               {
                  const int c = L.peers[pi].send_cells[k][j];
                  std::vector<int> &v = per_tile[tile_of[c]];
                  v.push_back (c);
                  v.push_back ((int) pi);
                  v.push_back (L.peers[pi].dst_start[k] + (int) j);
               }
         L.send_entries.clear ();
         L.n_send_tiles = 0;
         const int nh = L.tile_halo_max;
         for (int t = 0; t < L.n_tiles; ++t)
         {
            int *d = &L.rowdesc[(size_t) t * L.rowdesc_stride];
            int reads_ghost = 0;
            for (int i = 0; i < d[2]; ++i)
               if (d[rowd_off_halo () + i] >= L.n_owned) reads_ghost = 1;
            (void) nh;
            d[5] = reads_ghost;
            d[6] = (int) (L.send_entries.size () / 3);
            d[7] = (int) (per_tile[t].size () / 3);
            if (d[7]) ++L.n_send_tiles;
            L.send_entries.insert (L.send_entries.end (), per_tile[t].begin (), per_tile[t].end ());
         }
      }
      return true;
   }
}
