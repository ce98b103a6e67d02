#include "mesh.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <functional>
#include <map>
#include <sstream>
#include <unordered_map>

namespace dflo
{
   dflo_flat_mesh FlatMesh::view () const
   {
      dflo_flat_mesh m;
      m.n_cells = n_cells ();
      m.cell_origin = origin.data ();
      m.cell_size = size.data ();
      m.neighbor = neighbor.data ();
      m.face_flags = face_flags.data ();
      m.n_boundary_faces = n_bfaces ();
      m.bface_cell = bface_cell.data ();
      m.bface_face = bface_face.data ();
      m.bface_id = bface_id.data ();
      m.cell_vertices = vertices.data ();
      m.neighbor_face = neighbor_face.data ();
      m.n_hanging_faces = (int32_t) (hanging.size () / 6);
      m.hanging = hanging.data ();
      return m;
   }

   namespace
   {
      // local vertex pairs of the faces left, right, bottom, top (deal.II faces 0..3)
      const int FV[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};

      inline uint64_t edge_key (int a, int b)
      {
         const uint64_t lo = a < b ? a : b, hi = a < b ? b : a;
         return (hi << 32) | lo;
      }
   }

   bool flatten (const PrimitiveMesh &pm, const int bc_kind[DFLO_MAX_BOUNDARIES],
                 const int periodic_pair[DFLO_MAX_BOUNDARIES], FlatMesh &out, std::string &err)
   {
      const int nc = pm.n_cells ();
      out = FlatMesh ();
      out.origin.resize (2 * (size_t) nc);
      out.size.resize (2 * (size_t) nc);
      out.neighbor.assign (4 * (size_t) nc, -1);
      out.face_flags.assign (4 * (size_t) nc, 0);
      out.vertices.resize (8 * (size_t) nc);
      out.neighbor_face.resize (4 * (size_t) nc);
      for (size_t i = 0; i < out.neighbor_face.size (); ++i) out.neighbor_face[i] = (uint8_t) ((i & 3) ^ 1);

      struct Side { int cell, face; };
      std::unordered_map<uint64_t, std::pair<Side, Side>> edges;
      edges.reserve (2 * (size_t) nc + 16);
      const double *V = pm.vertices.data ();
      for (int c = 0; c < nc; ++c)
      {
         const int *v = &pm.cells[4 * (size_t) c];
         const double x0 = V[2 * v[0]], y0 = V[2 * v[0] + 1];
         const double hx = V[2 * v[1]] - x0, hy = V[2 * v[2] + 1] - y0;
         const double tol = 1e-12 * (std::fabs (hx) + std::fabs (hy));
         for (int i = 0; i < 4; ++i)
         {
            out.vertices[8 * (size_t) c + 2 * i] = V[2 * v[i]];
            out.vertices[8 * (size_t) c + 2 * i + 1] = V[2 * v[i] + 1];
         }
         if (!(hx > 0 && hy > 0) || std::fabs (V[2 * v[1] + 1] - y0) > tol || std::fabs (V[2 * v[2]] - x0) > tol
             || std::fabs (V[2 * v[3]] - (x0 + hx)) > tol || std::fabs (V[2 * v[3] + 1] - (y0 + hy)) > tol)
         {
            if (out.cartesian)
               out.why_not_cartesian = "cell " + std::to_string (c) + " is not an axis-aligned rectangle with lexicographic vertices (mapping = cartesian)";
            out.cartesian = false;
            // MappingQ1: the Jacobian of the bilinear map must be positive at the four vertices
            const double *q = &out.vertices[8 * (size_t) c];
            for (int i = 0; i < 4; ++i)
            {
               const double xi = i & 1, eta = i >> 1;
               const double xxi = (q[2] - q[0]) * (1 - eta) + (q[6] - q[4]) * eta, xeta = (q[4] - q[0]) * (1 - xi) + (q[6] - q[2]) * xi;
               const double yxi = (q[3] - q[1]) * (1 - eta) + (q[7] - q[5]) * eta, yeta = (q[5] - q[1]) * (1 - xi) + (q[7] - q[3]) * xi;
               if (!(xxi * yeta - xeta * yxi > 0.0))
               {
                  err = "cell " + std::to_string (c) + " has a non-positive Jacobian (vertices not in deal.II's lexicographic order, or not convex)";
                  return false;
               }
            }
         }
         out.origin[2 * (size_t) c] = x0;
         out.origin[2 * (size_t) c + 1] = y0;
         // off Cartesian meshes these are only nominal sizes (tile ordering heuristics); the q1 kernels use the vertices
         out.size[2 * (size_t) c] = out.cartesian ? hx : std::hypot (hx, V[2 * v[1] + 1] - y0);
         out.size[2 * (size_t) c + 1] = out.cartesian ? hy : std::hypot (V[2 * v[2]] - x0, hy);
         for (int f = 0; f < 4; ++f)
         {
            const uint64_t k = edge_key (v[FV[f][0]], v[FV[f][1]]);
            auto it = edges.find (k);
            if (it == edges.end ())
               edges.emplace (k, std::make_pair (Side{c, f}, Side{-1, -1}));
            else if (it->second.second.cell < 0)
               it->second.second = Side{c, f};
            else
            {
               err = "face shared by more than two cells";
               return false;
            }
         }
      }
      // hanging nodes: an unmatched edge A-B whose mid point M is a vertex, with the unmatched edges A-M and M-B on the other side
      {
         std::map<std::pair<long long, long long>, int> vertex_at;
         auto vkey = [] (double x, double y) { return std::make_pair ((long long) std::llround (x * 1e9), (long long) std::llround (y * 1e9)); };
         for (int v = 0; v < pm.n_vertices (); ++v) vertex_at[vkey (V[2 * v], V[2 * v + 1])] = v;
         std::vector<uint64_t> single;
         for (auto &kv : edges)
            if (kv.second.second.cell < 0) single.push_back (kv.first);
         std::sort (single.begin (), single.end ()); // a reproducible order of the table
         for (uint64_t k : single)
         {
            const Side c = edges[k].first;
            if (c.cell < 0) continue; // consumed as a child below
            const int A = pm.cells[4 * (size_t) c.cell + FV[c.face][0]], B = pm.cells[4 * (size_t) c.cell + FV[c.face][1]];
            auto it = vertex_at.find (vkey (0.5 * (V[2 * A] + V[2 * B]), 0.5 * (V[2 * A + 1] + V[2 * B + 1])));
            if (it == vertex_at.end ()) continue;
            const int M = it->second;
            auto e0 = edges.find (edge_key (A, M)), e1 = edges.find (edge_key (M, B));
            if (e0 == edges.end () || e1 == edges.end () || e0->second.second.cell >= 0 || e1->second.second.cell >= 0) continue;
            const Side kid[2] = {e0->second.first, e1->second.first};
            out.hanging.push_back (c.cell);
            out.hanging.push_back (c.face);
            for (int j = 0; j < 2; ++j)
            {
               out.hanging.push_back (kid[j].cell);
               out.hanging.push_back (kid[j].face);
               const bool rev = pm.cells[4 * (size_t) kid[j].cell + FV[kid[j].face][0]] != (j == 0 ? A : M);
               out.neighbor[4 * (size_t) kid[j].cell + kid[j].face] = c.cell;
               out.neighbor_face[4 * (size_t) kid[j].cell + kid[j].face] = (uint8_t) c.face;
               out.face_flags[4 * (size_t) kid[j].cell + kid[j].face] = DFLO_FACE_COARSER | (j ? DFLO_FACE_CHILD1 : 0) | (rev ? DFLO_FACE_FLIP : 0) | DFLO_FACE_OWNER;
            }
            out.neighbor[4 * (size_t) c.cell + c.face] = kid[0].cell;
            out.neighbor_face[4 * (size_t) c.cell + c.face] = (uint8_t) kid[0].face;
            out.face_flags[4 * (size_t) c.cell + c.face] = DFLO_FACE_HANGING;
            // the three edges are interior now: take them out of the boundary pass below
            edges[k].second = Side{-2, -2};
            e0->second.second = Side{-2, -2};
            e1->second.second = Side{-2, -2};
         }
      }
      std::unordered_map<uint64_t, int> line_id;
      for (int b = 0; b < pm.n_blines (); ++b) line_id[edge_key (pm.blines[2 * b], pm.blines[2 * b + 1])] = pm.bline_id[b];

      // boundary id of every (cell,face) that has no vertex-sharing neighbour; -1 elsewhere
      std::vector<int> bid (4 * (size_t) nc, -1);
      for (auto &kv : edges)
      {
         const Side a = kv.second.first, b = kv.second.second;
         if (b.cell == -2) continue; // a face with a hanging node, or one of its halves: handled above
         if (b.cell >= 0)
         {
            // do the two cells run along the shared line in the same direction?  (deal.II orients 2-D meshes so that they
            // do; a mesh that does not is still integrated correctly with the face points of one side reversed)
            const bool rev = pm.cells[4 * (size_t) a.cell + FV[a.face][0]] != pm.cells[4 * (size_t) b.cell + FV[b.face][0]];
            if ((a.face ^ 1) != b.face || rev)
            {
               if (out.cartesian) out.why_not_cartesian = "neighbouring cells are not equally oriented (mapping = cartesian)";
               out.cartesian = false;
            }
            out.neighbor[4 * (size_t) a.cell + a.face] = b.cell;
            out.neighbor[4 * (size_t) b.cell + b.face] = a.cell;
            out.neighbor_face[4 * (size_t) a.cell + a.face] = (uint8_t) b.face;
            out.neighbor_face[4 * (size_t) b.cell + b.face] = (uint8_t) a.face;
            // MeshWorker::loop integrates an interior face once, from the smaller cell
            out.face_flags[4 * (size_t) a.cell + a.face] = (a.cell < b.cell ? DFLO_FACE_OWNER : 0) | (rev ? DFLO_FACE_FLIP : 0);
            out.face_flags[4 * (size_t) b.cell + b.face] = (b.cell < a.cell ? DFLO_FACE_OWNER : 0) | (rev ? DFLO_FACE_FLIP : 0);
         }
         else
         {
            auto it = line_id.find (kv.first);
            const int id = it == line_id.end () ? 0 : it->second; // unlisted faces get id 0
            if (id < 0 || id >= DFLO_MAX_BOUNDARIES)
            {
               err = "boundary id " + std::to_string (id) + " out of range";
               return false;
            }
            bid[4 * (size_t) a.cell + a.face] = id;
         }
      }

      // periodic partners: same tangential coordinate, opposite face, partner id
      std::map<int, std::vector<Side>> by_id;
      for (int c = 0; c < nc; ++c)
         for (int f = 0; f < 4; ++f)
         {
            const int id = bid[4 * (size_t) c + f];
            if (id >= 0 && bc_kind[id] == DFLO_BC_PERIODIC) by_id[id].push_back (Side{c, f});
         }
      // A periodic boundary is a straight side of the domain: vertical (its faces share x) or horizontal.  Faces are
      // matched by the coordinate along the side; whether the two cells run along the line in the same direction is
      // read off their vertex order (DFLO_FACE_FLIP), and the partner's local face number is kept (neighbor_face) --
      // on lattice meshes that is the opposite face and no flip.
      auto face_end = [&] (const Side &s, int which, int d) { return out.vertices[8 * (size_t) s.cell + 2 * FV[s.face][which] + d]; };
      auto vertical = [&] (const std::vector<Side> &v) {
         double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
         for (const Side &s : v)
            for (int w = 0; w < 2; ++w)
               for (int d = 0; d < 2; ++d)
               {
                  lo[d] = std::min (lo[d], face_end (s, w, d));
                  hi[d] = std::max (hi[d], face_end (s, w, d));
               }
         return hi[0] - lo[0] < hi[1] - lo[1];
      };
      for (auto &kv : by_id)
      {
         const int partner = periodic_pair[kv.first];
         if (partner < 0 || partner >= DFLO_MAX_BOUNDARIES || !by_id.count (partner))
         {
            err = "periodic boundary " + std::to_string (kv.first) + " has no partner boundary";
            return false;
         }
         const int d = vertical (kv.second) ? 1 : 0; // coordinate along the side
         auto tangential = [&] (const Side &s) { return 0.5 * (face_end (s, 0, d) + face_end (s, 1, d)); };
         std::vector<Side> cand = by_id[partner];
         std::sort (cand.begin (), cand.end (), [&] (const Side &a, const Side &b) { return tangential (a) < tangential (b); });
         for (const Side &s : kv.second)
         {
            const double t = tangential (s);
            const double tol = 1e-9 * std::fabs (face_end (s, 1, d) - face_end (s, 0, d));
            auto lo = std::lower_bound (cand.begin (), cand.end (), t - tol, [&] (const Side &a, double val) { return tangential (a) < val; });
            if (lo == cand.end () || tangential (*lo) > t + tol)
            {
               err = "periodic face without partner";
               return false;
            }
            const bool rev = (face_end (s, 1, d) - face_end (s, 0, d)) * (face_end (*lo, 1, d) - face_end (*lo, 0, d)) < 0.0;
            if ((lo->face != (s.face ^ 1) || rev) && out.cartesian)
            {
               out.cartesian = false;
               out.why_not_cartesian = "periodic partners are not equally oriented (mapping = cartesian)";
            }
            out.neighbor[4 * (size_t) s.cell + s.face] = lo->cell;
            out.neighbor_face[4 * (size_t) s.cell + s.face] = (uint8_t) lo->face;
            out.face_flags[4 * (size_t) s.cell + s.face] = DFLO_FACE_PERIODIC | (rev ? DFLO_FACE_FLIP : 0);
         }
      }

      // genuine boundary faces ordered by (cell, face)
      for (int c = 0; c < nc; ++c)
         for (int f = 0; f < 4; ++f)
            if (out.neighbor[4 * (size_t) c + f] < 0)
            {
               out.neighbor[4 * (size_t) c + f] = -1 - (int) out.bface_cell.size ();
               out.bface_cell.push_back (c);
               out.bface_face.push_back (f);
               out.bface_id.push_back (bid[4 * (size_t) c + f]);
            }
      return true;
   }

   namespace
   {
      // Union of lattice-aligned blocks sharing one vertex lattice; boundary ids assigned to the
      // outer edges by a classifier on the edge mid point.
      struct LatticeBuilder
      {
         double X0, Y0, dx, dy;
         std::map<std::pair<int, int>, int> vid;
         PrimitiveMesh pm;

         int vertex (int i, int j)
         {
            auto key = std::make_pair (i, j);
            auto it = vid.find (key);
            if (it != vid.end ()) return it->second;
            const int id = pm.n_vertices ();
            pm.vertices.push_back (X0 + i * dx);
            pm.vertices.push_back (Y0 + j * dy);
            vid[key] = id;
            return id;
         }

         void block (int i0, int j0, int ni, int nj)
         {
            for (int j = 0; j < nj; ++j)
               for (int i = 0; i < ni; ++i)
               {
                  pm.cells.push_back (vertex (i0 + i, j0 + j));
                  pm.cells.push_back (vertex (i0 + i + 1, j0 + j));
                  pm.cells.push_back (vertex (i0 + i, j0 + j + 1));
                  pm.cells.push_back (vertex (i0 + i + 1, j0 + j + 1));
               }
         }

         // classify(xm, ym, face) -> boundary id
         void boundary (const std::function<int (double, double, int)> &classify)
         {
            std::unordered_map<uint64_t, std::pair<int, int>> cnt; // edge -> (count, cell*4+face)
            const int nc = pm.n_cells ();
            for (int c = 0; c < nc; ++c)
               for (int f = 0; f < 4; ++f)
               {
                  auto &e = cnt[edge_key (pm.cells[4 * c + FV[f][0]], pm.cells[4 * c + FV[f][1]])];
                  e.first++;
                  e.second = 4 * c + f;
               }
            std::vector<int> outer;
            for (auto &kv : cnt)
               if (kv.second.first == 1) outer.push_back (kv.second.second);
            std::sort (outer.begin (), outer.end ());
            for (int cf : outer)
            {
               const int c = cf / 4, f = cf % 4;
               const int a = pm.cells[4 * c + FV[f][0]], b = pm.cells[4 * c + FV[f][1]];
               const double xm = 0.5 * (pm.vertices[2 * a] + pm.vertices[2 * b]);
               const double ym = 0.5 * (pm.vertices[2 * a + 1] + pm.vertices[2 * b + 1]);
               pm.blines.push_back (a);
               pm.blines.push_back (b);
               pm.bline_id.push_back (classify (xm, ym, f));
            }
         }
      };
   }

   PrimitiveMesh make_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4])
   {
      LatticeBuilder lb;
      lb.X0 = x0;
      lb.Y0 = y0;
      lb.dx = (x1 - x0) / nx;
      lb.dy = (y1 - y0) / ny;
      lb.block (0, 0, nx, ny);
      lb.boundary ([&] (double, double, int f) { return ids[f]; });
      return lb.pm;
   }

   // A rectangle whose interior vertices are displaced smoothly (the boundary stays put, so periodic pairs and boundary
   // ids are those of the rectangle): straight-sided general quadrilaterals for mapping = q1.  rotate != 0 also turns
   // the vertex order of every other cell by 90 / 180 / 270 degrees, so that neighbours meet on arbitrary local faces
   // and run along shared lines in opposite directions -- what an unoriented gmsh file can look like.
   PrimitiveMesh make_skewed_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4], double amp, int rotate)
   {
      PrimitiveMesh pm = make_rectangle (nx, ny, x0, x1, y0, y1, ids);
      const double Lx = x1 - x0, Ly = y1 - y0, pi = 3.14159265358979323846;
      for (int v = 0; v < pm.n_vertices (); ++v)
      {
         const double X = pm.vertices[2 * v], Y = pm.vertices[2 * v + 1];
         const double s = std::sin (2.0 * pi * (X - x0) / Lx) * std::sin (2.0 * pi * (Y - y0) / Ly);
         pm.vertices[2 * v] = X + amp * Lx / (2.0 * pi) * 0.9 * s;
         pm.vertices[2 * v + 1] = Y + amp * Ly / (2.0 * pi) * 0.7 * s;
      }
      if (rotate)
         for (int c = 0; c < pm.n_cells (); ++c)
         {
            int *q = &pm.cells[4 * (size_t) c];
            for (int r = (c * 7 + c / nx) % 4; r > 0; --r) // one quarter turn: (v0 v1 v2 v3) -> (v1 v3 v0 v2), Jacobian stays positive
            {
               const int t[4] = {q[1], q[3], q[0], q[2]};
               for (int i = 0; i < 4; ++i) q[i] = t[i];
            }
         }
      return pm;
   }

   // nx x ny cells of [x0,x1] x [y0,y1]; the cells i0 <= i < i1, j0 <= j < j1 are replaced by their four children (one level,
   // what deal.II's refine_grid leaves behind): hanging nodes all around the patch.  Vertices live on the half-cell lattice.
   PrimitiveMesh make_refined_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4], int i0, int i1, int j0, int j1,
                                         int rotate)
   {
      PrimitiveMesh pm;
      const double hx = (x1 - x0) / nx, hy = (y1 - y0) / ny;
      std::map<std::pair<int, int>, int> vid;
      auto v = [&] (int i2, int j2) {
         auto key = std::make_pair (i2, j2);
         auto it = vid.find (key);
         if (it != vid.end ()) return it->second;
         const int id = pm.n_vertices ();
         pm.vertices.push_back (x0 + 0.5 * hx * i2);
         pm.vertices.push_back (y0 + 0.5 * hy * j2);
         vid[key] = id;
         return id;
      };
      auto quad = [&] (int a, int b, int s) {
         pm.cells.push_back (v (a, b));
         pm.cells.push_back (v (a + s, b));
         pm.cells.push_back (v (a, b + s));
         pm.cells.push_back (v (a + s, b + s));
      };
      for (int j = 0; j < ny; ++j)
         for (int i = 0; i < nx; ++i)
            if (i >= i0 && i < i1 && j >= j0 && j < j1)
               for (int dj = 0; dj < 2; ++dj)
                  for (int di = 0; di < 2; ++di) quad (2 * i + di, 2 * j + dj, 1);
            else
               quad (2 * i, 2 * j, 2);
      auto line = [&] (int a, int b, int id) {
         pm.blines.push_back (a);
         pm.blines.push_back (b);
         pm.bline_id.push_back (id);
      };
      // boundary lines at the resolution of the cells behind them
      for (int j = 0; j < ny; ++j)
         for (int side = 0; side < 2; ++side)
         {
            const int i = side ? nx - 1 : 0, i2 = side ? 2 * nx : 0;
            if (i >= i0 && i < i1 && j >= j0 && j < j1)
            {
               line (v (i2, 2 * j), v (i2, 2 * j + 1), ids[side]);
               line (v (i2, 2 * j + 1), v (i2, 2 * j + 2), ids[side]);
            }
            else
               line (v (i2, 2 * j), v (i2, 2 * j + 2), ids[side]);
         }
      for (int i = 0; i < nx; ++i)
         for (int side = 0; side < 2; ++side)
         {
            const int j = side ? ny - 1 : 0, j2 = side ? 2 * ny : 0;
            if (i >= i0 && i < i1 && j >= j0 && j < j1)
            {
               line (v (2 * i, j2), v (2 * i + 1, j2), ids[2 + side]);
               line (v (2 * i + 1, j2), v (2 * i + 2, j2), ids[2 + side]);
            }
            else
               line (v (2 * i, j2), v (2 * i + 2, j2), ids[2 + side]);
         }
      if (rotate) // mixed cell orientations (mapping = q1 only): coarse and fine cells meet on arbitrary faces, in either direction
         for (int c = 0; c < pm.n_cells (); ++c)
         {
            int *q = &pm.cells[4 * (size_t) c];
            for (int r = (c * 5 + c / 3) % 4; r > 0; --r)
            {
               const int t[4] = {q[1], q[3], q[0], q[2]};
               for (int i = 0; i < 4; ++i) q[i] = t[i];
            }
         }
      return pm;
   }

   // examples/compression_corner/corner.geo: a channel of height 3 whose floor turns up by 9.5 degrees at x = 1: two
   // transfinite blocks, [0,1] x [0,3] (nx1 x ny cells) and the ramp part up to x = 5 (nx2 x ny cells, vertical grid lines,
   // uniform in y between the ramp and the ceiling) -- trapezoids, i.e. mapping = q1.  Physical Line 1 walls (floor, ramp,
   // ceiling), 2 inflow (left), 3 outflow (right).  (the .geo counts points: n1 = 10, n2 = 30, n3 = 20 <-> 9, 29, 19 cells)
   PrimitiveMesh make_compression_corner (int nx1, int nx2, int ny)
   {
      const double H = 3.0, L1 = 1.0, L2 = 4.0, tn = std::tan (9.5 * 3.14159265358979323846 / 180.0);
      PrimitiveMesh pm;
      const int nx = nx1 + nx2;
      auto vid = [&] (int i, int j) { return j * (nx + 1) + i; };
      for (int j = 0; j <= ny; ++j)
         for (int i = 0; i <= nx; ++i)
         {
            const double x = i <= nx1 ? L1 * i / nx1 : L1 + L2 * (i - nx1) / nx2;
            const double yb = i <= nx1 ? 0.0 : tn * (x - L1);
            pm.vertices.push_back (x);
            pm.vertices.push_back (yb + (H - yb) * j / ny);
         }
      for (int blk = 0; blk < 2; ++blk) // cells block by block, like gmsh writes the two surfaces
         for (int j = 0; j < ny; ++j)
            for (int i = blk ? nx1 : 0; i < (blk ? nx : nx1); ++i)
            {
               pm.cells.push_back (vid (i, j));
               pm.cells.push_back (vid (i + 1, j));
               pm.cells.push_back (vid (i, j + 1));
               pm.cells.push_back (vid (i + 1, j + 1));
            }
      auto line = [&] (int a, int b, int id) {
         pm.blines.push_back (a);
         pm.blines.push_back (b);
         pm.bline_id.push_back (id);
      };
      for (int i = 0; i < nx; ++i)
      {
         line (vid (i, 0), vid (i + 1, 0), 1);
         line (vid (i, ny), vid (i + 1, ny), 1);
      }
      for (int j = 0; j < ny; ++j)
      {
         line (vid (0, j), vid (0, j + 1), 2);
         line (vid (nx, j), vid (nx, j + 1), 3);
      }
      return pm;
   }

   // examples/isentropic_vortex/grid.geo: [-5,5]^2; Physical Line 1 bottom, 2 right, 3 top, 4 left
   PrimitiveMesh make_isentropic_vortex_grid (int n)
   {
      const int ids[4] = {4, 2, 1, 3};
      return make_rectangle (n, n, -5.0, 5.0, -5.0, 5.0, ids);
   }

   // examples/sod_shock_tube/tube.geo: [0,1] x [0, ny*dx]; Physical Line 0 walls, 1 outlet (right),
   // 2 inlet (left)
   PrimitiveMesh make_sod_tube (int nx, int ny)
   {
      const double dx = 1.0 / nx;
      const int ids[4] = {2, 1, 0, 0};
      return make_rectangle (nx, ny, 0.0, 1.0, 0.0, dx * ny, ids);
   }

   // examples/double_mach_reflection/grid.geo with ny-1 := ny_cells: two transfinite blocks that
   // meet at x0 = 1/6 so that the wall start is a mesh line; ids: 0 bottom x<x0, 1 bottom x>x0
   // (the reflecting wall), 2 right, 3 top, 4 left
   PrimitiveMesh make_double_mach_grid (int ny_cells)
   {
      const double Lx = 4.0, Ly = 1.0, xs = 1.0 / 6.0;
      const double dy = Ly / ny_cells;
      const int n1 = (int) std::ceil (xs / dy - 1e-12);
      const int n2 = (int) std::ceil ((Lx - xs) / dy - 1e-12);
      LatticeBuilder lb;
      lb.X0 = xs - n1 * dy;
      lb.Y0 = 0.0;
      lb.dx = dy;
      lb.dy = dy;
      lb.block (0, 0, n1, ny_cells);
      lb.block (n1, 0, n2, ny_cells);
      const double xsplit = xs;
      lb.boundary ([&] (double xm, double, int f) {
         if (f == 0) return 4;
         if (f == 1) return 2;
         if (f == 3) return 3;
         return xm < xsplit ? 0 : 1;
      });
      return lb.pm;
   }

   // examples/forward_step/step.geo: three transfinite blocks of an L-shaped channel
   // [0,3]x[0,1] with a step of height 0.2 starting at x = 0.6; ids: 1 inflow (left), 2 walls,
   // 3 outlet (right)
   PrimitiveMesh make_forward_step_grid (double cl)
   {
      const int n1 = (int) std::lround (0.6 / cl), n2 = (int) std::lround (0.2 / cl), n3 = (int) std::lround (0.8 / cl),
                n4 = (int) std::lround (2.4 / cl);
      LatticeBuilder lb;
      lb.X0 = 0.0;
      lb.Y0 = 0.0;
      lb.dx = cl;
      lb.dy = cl;
      lb.block (0, 0, n1, n2);       // surface 1
      lb.block (0, n2, n1, n3);      // surface 2
      lb.block (n1, n2, n4, n3);     // surface 3
      const double xr = cl * (n1 + n4);
      lb.boundary ([&] (double xm, double, int f) {
         if (f == 0 && xm < 0.5 * cl) return 1;
         if (f == 1 && xm > xr - 0.5 * cl) return 3;
         return 2;
      });
      return lb.pm;
   }

   bool write_gmsh2 (const std::string &path, const PrimitiveMesh &pm)
   {
      FILE *fp = std::fopen (path.c_str (), "w");
      if (!fp) return false;
      std::fprintf (fp, "$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n", pm.n_vertices ());
      for (int v = 0; v < pm.n_vertices (); ++v)
         std::fprintf (fp, "%d %.17g %.17g 0\n", v + 1, pm.vertices[2 * v], pm.vertices[2 * v + 1]);
      std::fprintf (fp, "$EndNodes\n$Elements\n%d\n", pm.n_blines () + pm.n_cells ());
      int e = 1;
      for (int b = 0; b < pm.n_blines (); ++b, ++e)
         std::fprintf (fp, "%d 1 2 %d %d %d %d\n", e, pm.bline_id[b], pm.bline_id[b], pm.blines[2 * b] + 1, pm.blines[2 * b + 1] + 1);
      for (int c = 0; c < pm.n_cells (); ++c, ++e) // gmsh quads are counter-clockwise
         std::fprintf (fp, "%d 3 2 100 1 %d %d %d %d\n", e, pm.cells[4 * c] + 1, pm.cells[4 * c + 1] + 1, pm.cells[4 * c + 3] + 1,
                       pm.cells[4 * c + 2] + 1);
      std::fprintf (fp, "$EndElements\n");
      std::fclose (fp);
      return true;
   }

   bool read_gmsh2 (const std::string &path, PrimitiveMesh &pm, std::string &err)
   {
      std::ifstream in (path.c_str ());
      if (!in)
      {
         err = "cannot open " + path;
         return false;
      }
      pm = PrimitiveMesh ();
      std::string line;
      std::map<int, int> node_index;
      while (std::getline (in, line))
      {
         if (line.compare (0, 6, "$Nodes") == 0)
         {
            int n;
            in >> n;
            for (int i = 0; i < n; ++i)
            {
               int id;
               double x, y, z;
               in >> id >> x >> y >> z;
               node_index[id] = i;
               pm.vertices.push_back (x);
               pm.vertices.push_back (y);
            }
         }
         else if (line.compare (0, 9, "$Elements") == 0)
         {
            int n;
            in >> n;
            for (int i = 0; i < n; ++i)
            {
               int id, type, ntags;
               in >> id >> type >> ntags;
               std::vector<int> tags (ntags);
               for (int t = 0; t < ntags; ++t) in >> tags[t];
               const int nn = type == 1 ? 2 : type == 3 ? 4 : type == 15 ? 1 : -1;
               if (nn < 0)
               {
                  err = "unsupported gmsh element type " + std::to_string (type);
                  return false;
               }
               int nd[4];
               for (int k = 0; k < nn; ++k)
               {
                  in >> nd[k];
                  nd[k] = node_index[nd[k]];
               }
               if (type == 1)
               {
                  pm.blines.push_back (nd[0]);
                  pm.blines.push_back (nd[1]);
                  pm.bline_id.push_back (ntags > 0 ? tags[0] : 0); // boundary_id = physical tag
               }
               else if (type == 3)
               {
                  // counter-clockwise (or clockwise) ring -> lexicographic with v0 = lower-left
                  int r = 0;
                  for (int k = 1; k < 4; ++k)
                  {
                     const double dxk = pm.vertices[2 * nd[k]] - pm.vertices[2 * nd[r]];
                     const double dyk = pm.vertices[2 * nd[k] + 1] - pm.vertices[2 * nd[r] + 1];
                     if (dxk + dyk < 0) r = k;
                  }
                  int ring[4];
                  for (int k = 0; k < 4; ++k) ring[k] = nd[(r + k) % 4];
                  // ring[1] must be the +x neighbour; if it is the +y neighbour the ring is clockwise
                  const double dy1 = pm.vertices[2 * ring[1] + 1] - pm.vertices[2 * ring[0] + 1];
                  const double dx1 = pm.vertices[2 * ring[1]] - pm.vertices[2 * ring[0]];
                  if (std::fabs (dy1) > std::fabs (dx1)) std::swap (ring[1], ring[3]);
                  pm.cells.push_back (ring[0]);
                  pm.cells.push_back (ring[1]);
                  pm.cells.push_back (ring[3]);
                  pm.cells.push_back (ring[2]);
               }
            }
         }
      }
      if (pm.n_cells () == 0)
      {
         err = "no quadrilateral cells in " + path;
         return false;
      }
      return true;
   }
}
