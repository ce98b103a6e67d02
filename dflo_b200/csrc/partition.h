// Cell sharding of the flat mesh across ranks (one process per GPU).  Cells are split by global
// cell id into `world` contiguous ranges (what dflo's src_mpi tree gets from p4est, reference
// src_mpi/claw.cc:331-340, reduced to the structured case).  Each rank holds its owned cells plus
// `layers` layers of ghost cells in the face-neighbour graph:
//   layers = 1: the reference's one-layer halo, enough when no TVB limiter is active;
//   layers = 2: lets the stage kernel redundantly update the first ghost layer so that the TVB
//               limiter of owned cells sees its neighbours' post-update means without a second
//               exchange: still ONE halo exchange per RK stage.
// Local cell order: owned (global order) | ghost layer 1 grouped by owner rank | ghost layer 2
// grouped by owner rank, each group in global order, so every peer's data lands in at most two
// contiguous ranges and needs no unpack kernel.
#pragma once

#include "../../include/dflo_b200.h"

#include <algorithm>
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
      std::vector<int> bf_global, bf_id; // local boundary faces (of the computed cells) -> global bface, boundary id
      std::vector<int> bf_cell, bf_face; // local cell, face
      std::vector<HaloPeer> peers;
   };

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
      for (int c = b; c < e; ++c)
         for (int f = 0; f < 4; ++f)
         {
            const int nb = m.neighbor[4 * (size_t) c + f];
            if (nb >= 0 && (nb < b || nb >= e) && !mark[nb])
            {
               mark[nb] = 1;
               g1.push_back (nb);
            }
         }
      if (layers >= 2)
         for (int c : g1)
            for (int f = 0; f < 4; ++f)
            {
               const int nb = m.neighbor[4 * (size_t) c + f];
               if (nb >= 0 && (nb < b || nb >= e) && !mark[nb])
               {
                  mark[nb] = 2;
                  g2.push_back (nb);
               }
            }
      std::sort (g1.begin (), g1.end ()); // contiguous ownership => sorted by (owner, id)
      std::sort (g2.begin (), g2.end ());
   }

   inline bool build_local_mesh (const dflo_flat_mesh &m, int rank, int world, int layers, LocalMesh &L, std::string &err)
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
      for (int i = 0; i < L.n_owned; ++i) L.l2g[i] = L.begin + i;
      for (int i = 0; i < L.n_ghost1; ++i) L.l2g[L.n_owned + i] = g1[i];
      for (int i = 0; i < L.n_ghost2; ++i) L.l2g[L.n_owned + L.n_ghost1 + i] = g2[i];
      // global -> local lookup for the cells we hold
      auto g2l = [&] (int g) -> int {
         if (g >= L.begin && g < L.end) return g - L.begin;
         auto it = std::lower_bound (g1.begin (), g1.end (), g);
         if (it != g1.end () && *it == g) return L.n_owned + (int) (it - g1.begin ());
         it = std::lower_bound (g2.begin (), g2.end (), g);
         if (it != g2.end () && *it == g) return L.n_owned + L.n_ghost1 + (int) (it - g2.begin ());
         return -1;
      };
      L.nbr.resize (4 * (size_t) L.n_local);
      L.fflags.resize (4 * (size_t) L.n_local);
      L.geom.resize (4 * (size_t) L.n_local);
      for (int l = 0; l < L.n_local; ++l)
      {
         const int g = L.l2g[l];
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
      if (world == 1) return true;

      // Halo lists.  What we receive: our ghosts, grouped by owner.  What we send to peer p: the
      // cells of p's ghost layers that we own, in p's local order (global order).
      std::vector<HaloPeer> peers (world);
      for (int p = 0; p < world; ++p)
      {
         peers[p].rank = p;
         for (int k = 0; k < 2; ++k) peers[p].recv_start[k] = peers[p].recv_count[k] = 0;
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
         for (int g : pg1)
            if (g >= L.begin && g < L.end) peers[p].send_cells[0].push_back (g - L.begin);
         for (int g : pg2)
            if (g >= L.begin && g < L.end) peers[p].send_cells[1].push_back (g - L.begin);
      }
      for (int p = 0; p < world; ++p)
         if (p != rank
             && (peers[p].recv_count[0] || peers[p].recv_count[1] || !peers[p].send_cells[0].empty () || !peers[p].send_cells[1].empty ()))
            L.peers.push_back (peers[p]);
      return true;
   }
}
