// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product; the product never links this.
//
// CPU restatement of dflo's explicit DG path, structured per cell / per face exactly like the
// reference (dense i x q loops, one flux call per quadrature point through the flux_type
// switch), on top of a restatement of the deal.II pieces the path relies on (SURVEY.md
// Appendix A1-A7: unit cell & face numbering, QGauss/QGaussLobatto, FE_DGQArbitraryNodes on
// Gauss nodes, FE_DGP orthonormal Legendre, FESystem numbering, MappingCartesian, the
// MeshWorker::loop visiting rule and the ResidualSimple copier).  deal.II is a third-party
// dependency that is NOT under /root/reference (src/CMakeLists.txt:26 asks for >= 8.0,
// README.md:4 suggests 9.5.0); its published algorithms are restated here and anchored on the
// reference's own call sites, cited per function.
//
// PARITY PIN: physics (phys.h) is pinned bit-for-bit to the reference's equation.h object code;
// the assembly level is "parity unpinned" (the reference has no tests/golden vectors and cannot
// be built here) -- see DESIGN.md.
#include "dflo_oracle.h"
#include "phys.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <utility>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace
{
   std::string g_error;

   const int NC = 4;      // EulerEquations<2>::n_components, equation.h:26
   const int RHO = 2;     // density_component
   const int ENE = 3;     // energy_component

   //---------------------------------------------------------------------------------------------
   // 1-D quadrature rules on [0,1] (deal.II QGauss<1>, QGaussLobatto<1>)
   //---------------------------------------------------------------------------------------------
   void legendre (int n, double t, double &P, double &dP)
   {
      // P_n(t), P_n'(t) on [-1,1] by the three-term recurrence
      double p0 = 1.0, p1 = t;
      if (n == 0) { P = 1.0; dP = 0.0; return; }
      for (int j = 2; j <= n; ++j)
      {
         const double p2 = ((2.0 * j - 1.0) * t * p1 - (j - 1.0) * p0) / j;
         p0 = p1;
         p1 = p2;
      }
      P = p1;
      dP = n * (t * p1 - p0) / (t * t - 1.0);
   }

   void gauss_rule (int n, std::vector<double> &x, std::vector<double> &w)
   {
      x.resize (n);
      w.resize (n);
      for (int i = 0; i < n; ++i)
      {
         double t = -std::cos (M_PI * (i + 0.75) / (n + 0.5)); // ascending
         for (int it = 0; it < 100; ++it)
         {
            double P, dP;
            legendre (n, t, P, dP);
            const double dt = P / dP;
            t -= dt;
            if (std::fabs (dt) < 1e-16) break;
         }
         double P, dP;
         legendre (n, t, P, dP);
         x[i] = 0.5 * (1.0 + t);
         w[i] = 1.0 / ((1.0 - t * t) * dP * dP); // = (2/((1-t^2)P'^2))/2
      }
      // symmetrise (deal.II does the same: nodes are mirrored)
      for (int i = 0; i < n / 2; ++i)
      {
         const double xm = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
         x[i] = xm;
         x[n - 1 - i] = 1.0 - xm;
         const double wm = 0.5 * (w[i] + w[n - 1 - i]);
         w[i] = w[n - 1 - i] = wm;
      }
      if (n % 2 == 1) x[n / 2] = 0.5;
   }

   void gauss_lobatto_rule (int n, std::vector<double> &x)
   {
      // end points + roots of P'_{n-1}; only the nodes are needed (positivity.cc:44-47 uses
      // update_values only)
      x.assign (n, 0.0);
      x[0] = 0.0;
      x[n - 1] = 1.0;
      const int m = n - 1;
      for (int i = 1; i < n - 1; ++i)
      {
         double t = -std::cos (M_PI * i / m);
         for (int it = 0; it < 100; ++it)
         {
            double P, dP;
            legendre (m, t, P, dP);
            // d/dt of P'_m via Legendre ODE: (1-t^2) P'' = 2 t P' - m(m+1) P
            const double d2P = (2.0 * t * dP - m * (m + 1.0) * P) / (1.0 - t * t);
            const double dt = dP / d2P;
            t -= dt;
            if (std::fabs (dt) < 1e-16) break;
         }
         x[i] = 0.5 * (1.0 + t);
      }
      for (int i = 0; i < n / 2; ++i)
      {
         const double xm = 0.5 * (x[i] + (1.0 - x[n - 1 - i]));
         x[i] = xm;
         x[n - 1 - i] = 1.0 - xm;
      }
      if (n % 2 == 1) x[n / 2] = 0.5;
   }

   //---------------------------------------------------------------------------------------------
   // Scalar finite elements on the unit square
   //---------------------------------------------------------------------------------------------
   // Lagrange polynomial a on nodes xs (FE_DGQArbitraryNodes(QGauss<1>(k+1)), main.cc:40; A3)
   void lagrange (const std::vector<double> &xs, int a, double x, double &l, double &dl)
   {
      const int n = xs.size ();
      l = 1.0;
      for (int j = 0; j < n; ++j)
         if (j != a) l *= (x - xs[j]) / (xs[a] - xs[j]);
      dl = 0.0;
      for (int m = 0; m < n; ++m)
      {
         if (m == a) continue;
         double t = 1.0 / (xs[a] - xs[m]);
         for (int j = 0; j < n; ++j)
            if (j != a && j != m) t *= (x - xs[j]) / (xs[a] - xs[j]);
         dl += t;
      }
   }

   // Orthonormal Legendre on [0,1]: L_i(x) = sqrt(2i+1) P_i(2x-1) (FE_DGP, main.cc:46; A4;
   // the sqrt(3) of limiter.cc:395,417 is L_1's slope normalisation)
   void legendre01 (int i, double x, double &L, double &dL)
   {
      double P, dP;
      const double t = 2.0 * x - 1.0;
      if (std::fabs (std::fabs (t) - 1.0) < 1e-14)
      {
         // end points: P_i(+-1) = (+-1)^i, P_i'(+-1) = (+-1)^(i-1) i(i+1)/2
         const double s = t > 0 ? 1.0 : -1.0;
         P = (i % 2 == 0) ? 1.0 : s;
         dP = ((i % 2 == 1) ? 1.0 : s) * 0.5 * i * (i + 1.0);
      }
      else
         legendre (i, t, P, dP);
      const double nrm = std::sqrt (2.0 * i + 1.0);
      L = nrm * P;
      dL = nrm * 2.0 * dP;
   }

   struct FE
   {
      int basis, k, n1, ns, D;
      std::vector<double> gx, gw;       // QGauss<1>(k+1)
      std::vector<int> px, py;          // Pk: x/y degree of scalar basis m (claw.cc:104-114)

      void init (int basis_, int k_)
      {
         basis = basis_;
         k = k_;
         n1 = k + 1;
         gauss_rule (n1, gx, gw);
         if (basis == ORACLE_BASIS_QK)
            ns = n1 * n1;
         else
         {
            for (int j = 0; j <= k; ++j)
               for (int i = 0; i <= k - j; ++i)
               {
                  px.push_back (i);
                  py.push_back (j);
               }
            ns = px.size ();
         }
         D = NC * ns;
      }

      // value and unit-cell gradient of scalar basis function m at (x,y)
      void eval (int m, double x, double y, double &v, double g[2]) const
      {
         double lx, dlx, ly, dly;
         if (basis == ORACLE_BASIS_QK)
         {
            lagrange (gx, m % n1, x, lx, dlx); // DoF index = a + (k+1) b, x fastest (A3)
            lagrange (gx, m / n1, y, ly, dly);
         }
         else
         {
            legendre01 (px[m], x, lx, dlx);
            legendre01 (py[m], y, ly, dly);
         }
         v = lx * ly;
         g[0] = dlx * ly;
         g[1] = lx * dly;
      }
   };

   // An FEValues-like table on the unit cell: values/gradients of the ns scalar functions at a
   // list of points.  shape_value_component(i,q,c) of the FESystem is phi[(i%ns)][q] when
   // c == i/ns and zero otherwise (A5).
   struct UnitValues
   {
      int nq;
      std::vector<double> x, y, w;     // unit points and weights
      std::vector<double> phi, dphi;   // [ns][nq], [ns][nq][2]
      void init (const FE &fe, const std::vector<double> &x_, const std::vector<double> &y_,
                 const std::vector<double> &w_)
      {
         x = x_;
         y = y_;
         w = w_;
         nq = x.size ();
         phi.resize (fe.ns * nq);
         dphi.resize (fe.ns * nq * 2);
         for (int m = 0; m < fe.ns; ++m)
            for (int q = 0; q < nq; ++q)
               fe.eval (m, x[q], y[q], phi[m * nq + q], &dphi[(m * nq + q) * 2]);
      }
   };

   struct Cell
   {
      double x0, y0, hx, hy;
      double vx[4], vy[4]; // vertices, deal.II lexicographic order (A1); MappingQ1 works from these
      int nbr_face[4];   // the neighbour's local number of the shared face (f ^ 1 on lattice meshes)
      int sub[4];        // hanging nodes: >= 0: the neighbour across face f is COARSER and this face is its child `sub` (0: the half at
                         // the start of the coarse face's line, 1: the other); -1 otherwise
      int hang[4];       // >= 0: face f has a hanging node (two finer neighbours): index into oracle_ctx::hanging; -1 otherwise
      int nbr[4];        // neighbour cell or -1 (boundary)
      int bid[4];        // boundary id if at boundary
      int bface[4];      // index into the non-periodic boundary-face list, or -1
      bool flip[4];      // the neighbour runs through the face quadrature points in the opposite order (periodic face_flip;
                         // general quads whose two cells see the shared line in opposite directions)
   };

   const double NORMAL[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}}; // MappingCartesian, A6

   // a face with a hanging node: the coarse cell's face and the two finer cells behind it, in the order of the coarse line
   struct Hanging
   {
      int coarse, cface;
      int fine[2], fface[2];
   };
}

struct oracle_ctx
{
   oracle_params prm;
   FE fe;
   UnitValues vol;          // QGauss<2>(k+1), x fastest (claw.cc:419-422)
   UnitValues face[4];      // QGauss<1>(k+1) projected to face f (A2)
   UnitValues posx, posy;   // positivity.cc:43-47
   UnitValues support;      // Qk unit support points (limiter.cc:234)
   UnitValues mmgrad;       // QGauss<2>(nq) of the minmax limiter (src_mpi/limiter.cc:407-413)
   std::vector<double> ext_force; // [nc][nq][2] external force at the cell quadrature points (MPI tree), empty: src forcing
   std::vector<Cell> cells;
   std::vector<char> shared;    // [nc][4] face shared through vertices (cell->neighbor() exists)
   int n_bfaces;
   std::vector<int> bf_cell, bf_face, bf_bid;
   std::vector<double> bc_values;          // [nbf][nqf][4]

   std::vector<double> current, old, rhs, newton_update, dt;
   std::vector<double> cell_average;       // [nc][4]
   std::vector<double> inv_mass;           // [nc][D]  claw.cc:228-258
   std::vector<double> shock_indicator;
   std::vector<int> limited;
   double global_dt;
   double ark[3];
   int n_rk;

   UnitValues dtq;          // QIterated(QTrapez,3): the 4x4 equispaced points of compute_time_step_q (claw.cc:522)
   UnitValues subface[4][2]; // FESubfaceValues: QGauss<1>(k+1) on child c of face f of the COARSE cell (MeshWorker, hanging nodes)
   std::vector<Hanging> hanging;

   int D () const { return fe.D; }
   bool q1 () const { return prm.mapping == ORACLE_MAPPING_Q1; }
   // MappingQ1: x(xi,eta) = sum_v N_v(xi,eta) x_v, N = (1-xi)(1-eta), xi(1-eta), (1-xi)eta, xi eta;  J = d(x,y)/d(xi,eta)
   void jacobian (const Cell &c, double xi, double eta, double J[4]) const
   {
      J[0] = (c.vx[1] - c.vx[0]) * (1.0 - eta) + (c.vx[3] - c.vx[2]) * eta; // x_xi
      J[1] = (c.vx[2] - c.vx[0]) * (1.0 - xi) + (c.vx[3] - c.vx[1]) * xi;   // x_eta
      J[2] = (c.vy[1] - c.vy[0]) * (1.0 - eta) + (c.vy[3] - c.vy[2]) * eta; // y_xi
      J[3] = (c.vy[2] - c.vy[0]) * (1.0 - xi) + (c.vy[3] - c.vy[1]) * xi;   // y_eta
   }
   void map_point (const Cell &c, double xi, double eta, double &x, double &y) const
   {
      if (!q1 ())
      {
         x = c.x0 + xi * c.hx;
         y = c.y0 + eta * c.hy;
         return;
      }
      const double n0 = (1.0 - xi) * (1.0 - eta), n1 = xi * (1.0 - eta), n2 = (1.0 - xi) * eta, n3 = xi * eta;
      x = n0 * c.vx[0] + n1 * c.vx[1] + n2 * c.vx[2] + n3 * c.vx[3];
      y = n0 * c.vy[0] + n1 * c.vy[1] + n2 * c.vy[2] + n3 * c.vy[3];
   }
   double JxW (const Cell &c, int q) const
   {
      if (!q1 ()) return vol.w[q] * c.hx * c.hy;
      double J[4];
      jacobian (c, vol.x[q], vol.y[q], J);
      return vol.w[q] * (J[0] * J[3] - J[1] * J[2]);
   }
   // cell->diameter(): the longer diagonal;  cell->measure(): the area of the quadrilateral
   double diameter (const Cell &c) const
   {
      if (!q1 ()) return std::sqrt (c.hx * c.hx + c.hy * c.hy);
      const double d1 = std::hypot (c.vx[3] - c.vx[0], c.vy[3] - c.vy[0]), d2 = std::hypot (c.vx[2] - c.vx[1], c.vy[2] - c.vy[1]);
      return std::max (d1, d2);
   }
   double measure (const Cell &c) const
   {
      if (!q1 ()) return c.hx * c.hy;
      return 0.5 * ((c.vx[3] - c.vx[0]) * (c.vy[2] - c.vy[1]) - (c.vx[2] - c.vx[1]) * (c.vy[3] - c.vy[0]));
   }
   // outward unit normal and length of face f (straight edge from vertex FACE_VERT[f][0] to FACE_VERT[f][1])
   void face_geometry (const Cell &c, int f, double n[2], double &len) const
   {
      static const int FV[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
      static const double CART[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
      if (!q1 ())
      {
         n[0] = CART[f][0];
         n[1] = CART[f][1];
         len = f < 2 ? c.hy : c.hx;
         return;
      }
      const double tx = c.vx[FV[f][1]] - c.vx[FV[f][0]], ty = c.vy[FV[f][1]] - c.vy[FV[f][0]];
      len = std::sqrt (tx * tx + ty * ty);
      const double s = (f == 1 || f == 2) ? 1.0 : -1.0; // (t_y, -t_x) points out of faces 1 and 2 of a positively oriented cell
      n[0] = s * ty / len;
      n[1] = -s * tx / len;
   }
};

namespace
{
   //---------------------------------------------------------------------------------------------
   // Triangulation: neighbours from shared vertex pairs; boundary ids from line elements
   //---------------------------------------------------------------------------------------------
   const int FACE_VERT[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}}; // A1

   bool build_mesh (oracle_ctx &o, int nv, const double *V, int nc, const int *C, int nb,
                    const int *BL, const int *BID)
   {
      (void) nv;
      o.cells.resize (nc);
      typedef std::pair<int, int> Key;
      std::map<Key, std::vector<std::pair<int, int>>> face_map;
      auto key = [] (int a, int b) { return a < b ? Key (a, b) : Key (b, a); };
      for (int c = 0; c < nc; ++c)
      {
         const int *v = C + 4 * c;
         Cell &cl = o.cells[c];
         cl.x0 = V[2 * v[0]];
         cl.y0 = V[2 * v[0] + 1];
         cl.hx = V[2 * v[1]] - cl.x0;
         cl.hy = V[2 * v[2] + 1] - cl.y0;
         const double tol = 1e-12 * (std::fabs (cl.hx) + std::fabs (cl.hy));
         const bool ok = cl.hx > 0 && cl.hy > 0 && std::fabs (V[2 * v[1] + 1] - cl.y0) <= tol
                         && std::fabs (V[2 * v[2]] - cl.x0) <= tol
                         && std::fabs (V[2 * v[3]] - (cl.x0 + cl.hx)) <= tol
                         && std::fabs (V[2 * v[3] + 1] - (cl.y0 + cl.hy)) <= tol;
         for (int i = 0; i < 4; ++i)
         {
            cl.vx[i] = V[2 * v[i]];
            cl.vy[i] = V[2 * v[i] + 1];
         }
         if (!ok && !o.q1 ())
         {
            g_error = "cell is not an axis-aligned rectangle in lexicographic orientation (MappingCartesian)";
            return false;
         }
         if (o.q1 ())
         {
            // MappingQ1 needs a positive Jacobian at every vertex (convex, counter-clockwise in lexicographic order)
            for (int i = 0; i < 4; ++i)
            {
               double J[4];
               o.jacobian (cl, i & 1, i >> 1, J);
               if (!(J[0] * J[3] - J[1] * J[2] > 0.0))
               {
                  g_error = "cell with a non-positive Jacobian (vertex order is not deal.II's lexicographic one, or the cell is not convex)";
                  return false;
               }
            }
         }
         for (int f = 0; f < 4; ++f)
         {
            face_map[key (v[FACE_VERT[f][0]], v[FACE_VERT[f][1]])].push_back (std::make_pair (c, f));
            cl.nbr[f] = -1;
            cl.sub[f] = -1;
            cl.hang[f] = -1;
            cl.nbr_face[f] = f ^ 1;
            cl.bid[f] = 0;
            cl.bface[f] = -1;
            cl.flip[f] = false;
         }
      }
      std::map<Key, int> bmap;
      for (int b = 0; b < nb; ++b) bmap[key (BL[2 * b], BL[2 * b + 1])] = BID[b];
      for (auto &kv : face_map)
      {
         auto &lst = kv.second;
         if (lst.size () == 2)
         {
            o.cells[lst[0].first].nbr[lst[0].second] = lst[1].first;
            o.cells[lst[1].first].nbr[lst[1].second] = lst[0].first;
            o.cells[lst[0].first].nbr_face[lst[0].second] = lst[1].second;
            o.cells[lst[1].first].nbr_face[lst[1].second] = lst[0].second;
            // same two vertices; opposite order = the two cells run along the line in opposite directions
            const bool rev = C[4 * lst[0].first + FACE_VERT[lst[0].second][0]] != C[4 * lst[1].first + FACE_VERT[lst[1].second][0]];
            o.cells[lst[0].first].flip[lst[0].second] = o.cells[lst[1].first].flip[lst[1].second] = rev;
         }
         else if (lst.size () == 1)
         {
            auto it = bmap.find (kv.first);
            o.cells[lst[0].first].bid[lst[0].second] = (it == bmap.end ()) ? 0 : it->second;
         }
         else
         {
            g_error = "non-manifold face";
            return false;
         }
      }
      // hanging nodes (one level of refinement, as deal.II keeps it): an unmatched edge A-B whose mid point M is a vertex with
      // the two unmatched edges A-M and M-B on the other side
      {
         std::map<std::pair<long long, long long>, int> vertex_at;
         auto vkey = [] (double x, double y) { return std::make_pair ((long long) std::llround (x * 1e9), (long long) std::llround (y * 1e9)); };
         for (int v = 0; v < nv; ++v) vertex_at[vkey (V[2 * v], V[2 * v + 1])] = v;
         std::vector<std::pair<Key, std::pair<int, int>>> single;
         for (auto &kv : face_map)
            if (kv.second.size () == 1) single.push_back (std::make_pair (kv.first, kv.second[0]));
         for (auto &e : single)
         {
            const int c = e.second.first, f = e.second.second;
            const int A = C[4 * c + FACE_VERT[f][0]], B = C[4 * c + FACE_VERT[f][1]];
            auto it = vertex_at.find (vkey (0.5 * (V[2 * A] + V[2 * B]), 0.5 * (V[2 * A + 1] + V[2 * B + 1])));
            if (it == vertex_at.end ()) continue;
            const int M = it->second;
            auto e0 = face_map.find (key (A, M)), e1 = face_map.find (key (M, B));
            if (e0 == face_map.end () || e1 == face_map.end () || e0->second.size () != 1 || e1->second.size () != 1) continue;
            Hanging h;
            h.coarse = c;
            h.cface = f;
            const std::pair<int, int> kids[2] = {e0->second[0], e1->second[0]};
            for (int k = 0; k < 2; ++k)
            {
               const int fc = kids[k].first, ff = kids[k].second;
               h.fine[k] = fc;
               h.fface[k] = ff;
               Cell &fcl = o.cells[fc];
               fcl.nbr[ff] = c;
               fcl.nbr_face[ff] = f;
               fcl.sub[ff] = k;
               // same direction along the line if the fine face starts where its half of the coarse face starts
               fcl.flip[ff] = C[4 * fc + FACE_VERT[ff][0]] != (k == 0 ? A : M);
            }
            o.cells[c].hang[f] = (int) o.hanging.size ();
            o.cells[c].nbr[f] = h.fine[0]; // "has a neighbour" (cell->neighbor(f) exists and has children)
            o.hanging.push_back (h);
         }
      }
      // vertex-sharing interior faces = what deal.II's cell->neighbor() knows about
      o.shared.assign ((size_t) 4 * nc, 0);
      for (int c = 0; c < nc; ++c)
         for (int f = 0; f < 4; ++f) o.shared[4 * c + f] = o.cells[c].nbr[f] >= 0;
      // periodic pairs (src_mpi/claw.cc:156-204 via GridTools::collect_periodic_faces): a
      // boundary face with a periodic id is matched with the boundary face of the partner id
      // that has the opposite face number and the same tangential coordinate
      std::map<int, std::vector<std::pair<int, int>>> by_id;
      for (int c = 0; c < nc; ++c)
         for (int f = 0; f < 4; ++f)
            if (!o.shared[4 * c + f]) by_id[o.cells[c].bid[f]].push_back (std::make_pair (c, f));
      for (auto &kv : by_id)
      {
         const int id = kv.first;
         if (id < 0 || id >= 10 || o.prm.bc_kind[id] != ORACLE_BC_PERIODIC) continue;
         const std::vector<std::pair<int, int>> &partners = by_id[o.prm.periodic_pair[id]];
         // a periodic boundary is a straight side of the domain, vertical or horizontal: match faces by the coordinate
         // along it (GridTools::collect_periodic_faces matches by the face centres up to the offset between the sides)
         auto end = [&] (const std::pair<int, int> &cf, int which, int d) {
            const Cell &k = o.cells[cf.first];
            const int v = FACE_VERT[cf.second][which];
            return d == 0 ? k.vx[v] : k.vy[v];
         };
         double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
         for (auto &cf : kv.second)
            for (int w = 0; w < 2; ++w)
               for (int d = 0; d < 2; ++d)
               {
                  lo[d] = std::min (lo[d], end (cf, w, d));
                  hi[d] = std::max (hi[d], end (cf, w, d));
               }
         const int d = (hi[0] - lo[0] < hi[1] - lo[1]) ? 1 : 0;
         for (auto &cf : kv.second)
         {
            const int c = cf.first, f = cf.second;
            Cell &cl = o.cells[c];
            const double tc = 0.5 * (end (cf, 0, d) + end (cf, 1, d));
            int found = -1;
            for (auto &cf2 : partners)
            {
               const double tc2 = 0.5 * (end (cf2, 0, d) + end (cf2, 1, d));
               if (std::fabs (tc - tc2) < 1e-9 * std::fabs (end (cf, 1, d) - end (cf, 0, d)))
               {
                  found = cf2.first;
                  cl.nbr_face[f] = cf2.second;
                  cl.flip[f] = (end (cf, 1, d) - end (cf, 0, d)) * (end (cf2, 1, d) - end (cf2, 0, d)) < 0.0;
               }
            }
            if (found < 0)
            {
               g_error = "periodic face without partner";
               return false;
            }
            cl.nbr[f] = found;
         }
      }
      // list of genuine boundary faces ordered by (cell, face)
      o.n_bfaces = 0;
      for (int c = 0; c < nc; ++c)
         for (int f = 0; f < 4; ++f)
            if (o.cells[c].nbr[f] < 0)
            {
               o.cells[c].bface[f] = o.n_bfaces++;
               o.bf_cell.push_back (c);
               o.bf_face.push_back (f);
               o.bf_bid.push_back (o.cells[c].bid[f]);
            }
      return true;
   }

   //---------------------------------------------------------------------------------------------
   // claw.cc:228-258
   //---------------------------------------------------------------------------------------------
   void compute_inv_mass_matrix (oracle_ctx &o)
   {
      const int D = o.D (), ns = o.fe.ns, nq = o.vol.nq;
      o.inv_mass.resize (o.cells.size () * D);
      for (size_t c = 0; c < o.cells.size (); ++c)
         for (int i = 0; i < D; ++i)
         {
            double m = 0.0;
            for (int q = 0; q < nq; ++q)
               m += o.vol.phi[(i % ns) * nq + q] * o.vol.phi[(i % ns) * nq + q] * o.JxW (o.cells[c], q);
            o.inv_mass[c * D + i] = 1.0 / m;
         }
   }

   //---------------------------------------------------------------------------------------------
   // assemble_explicit.cc:30-120
   //---------------------------------------------------------------------------------------------
   void integrate_cell_term (const oracle_ctx &o, int cno, double *local)
   {
      const Cell &cl = o.cells[cno];
      const int D = o.D (), ns = o.fe.ns, nq = o.vol.nq;
      const double *u = &o.current[(size_t) cno * D];
      std::vector<double> W (nq * NC), flux (nq * 8), forcing (nq * NC);
      for (int q = 0; q < nq; ++q)
      {
         for (int c = 0; c < NC; ++c) W[q * NC + c] = 0.0;
         for (int i = 0; i < D; ++i)
         {
            const int c = i / ns;
            W[q * NC + c] += u[i] * o.vol.phi[(i % ns) * nq + q];
         }
         phys_flux_matrix (&W[q * NC], &flux[q * 8]);
         if (o.ext_force.empty ())
            phys_forcing (&W[q * NC], &forcing[q * NC]);                                           // src: equation.h:829-850
         else // src_mpi/assemble_explicit.cc:56-58, 84
            phys_ext_forcing_restated (&W[q * NC], &o.ext_force[((size_t) cno * nq + q) * 2], &forcing[q * NC]);
      }
      const double h[2] = {cl.hx, cl.hy};
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns, bi = i % ns;
         for (int q = 0; q < nq; ++q)
         {
            if (o.q1 ())
            {
               // shape_grad = J^-T grad_unit (MappingQ1 with update_gradients)
               double J[4];
               o.jacobian (cl, o.vol.x[q], o.vol.y[q], J);
               const double det = J[0] * J[3] - J[1] * J[2];
               const double gu = o.vol.dphi[(bi * nq + q) * 2], gv = o.vol.dphi[(bi * nq + q) * 2 + 1];
               const double g[2] = {(J[3] * gu - J[2] * gv) / det, (-J[1] * gu + J[0] * gv) / det};
               for (int d = 0; d < 2; ++d) F_i -= flux[q * 8 + 2 * ci + d] * g[d] * o.JxW (cl, q);
            }
            else
            for (int d = 0; d < 2; ++d)
               F_i -= flux[q * 8 + 2 * ci + d] * (o.vol.dphi[(bi * nq + q) * 2 + d] / h[d]) * o.JxW (cl, q);
            F_i -= o.prm.gravity * forcing[q * NC + ci] * o.vol.phi[bi * nq + q] * o.JxW (cl, q);
         }
         local[i] -= F_i;
      }
   }

   // trace of cell cno on its face f at all face quadrature points (the q x i loops of
   // assemble_explicit.cc:176-193 / 303-334)
   void face_values (const oracle_ctx &o, int cno, int f, double *W /*[nqf][4]*/)
   {
      const int D = o.D (), ns = o.fe.ns, nqf = o.face[f].nq;
      const double *u = &o.current[(size_t) cno * D];
      for (int q = 0; q < nqf; ++q)
      {
         for (int c = 0; c < NC; ++c) W[q * NC + c] = 0.0;
         for (int i = 0; i < D; ++i)
         {
            const int c = i / ns;
            W[q * NC + c] += u[i] * o.face[f].phi[(i % ns) * nqf + q];
         }
      }
   }

   double face_JxW (const oracle_ctx &o, const Cell &cl, int f, int q)
   {
      if (o.q1 ())
      {
         double n[2], len;
         o.face_geometry (cl, f, n, len);
         return o.face[f].w[q] * len;
      }
      return o.face[f].w[q] * (f < 2 ? cl.hy : cl.hx);
   }

   //---------------------------------------------------------------------------------------------
   // compute_shock_indicator, indicator.cc:15-31 and compute_shock_indicator_kxrcf, 50-198
   // (same-level neighbours only: the engine has no hanging nodes).  Periodic faces are boundary
   // faces to cell->at_boundary(): like the reference, nothing is done on them.
   //---------------------------------------------------------------------------------------------
   void compute_shock_indicator (oracle_ctx &o)
   {
      const int nc = o.cells.size ();
      if (o.prm.shock_indicator == 0)
      {
         std::fill (o.shock_indicator.begin (), o.shock_indicator.end (), 1e20); // :18-22
         return;
      }
      const int component = o.prm.shock_indicator == 1 ? RHO : ENE;               // :70-82
      const int nqf = o.fe.n1;
      std::vector<double> W (nqf * NC), Wn (nqf * NC);
      for (int c = 0; c < nc; ++c)
      {
         const Cell &cl = o.cells[c];
         double cell_shock_ind = 0, inflow_measure = 0;
         double vel[2];
         for (int i = 0; i < 2; ++i) vel[i] = o.cell_average[c * NC + i] / o.cell_average[c * NC + RHO]; // :108-110
         for (int f = 0; f < 4; ++f)
         {
            if (!o.shared[c * 4 + f]) continue;                                   // at_boundary: nothing, :181-186
            const int n = cl.nbr[f], nf = f ^ 1;                                  // neighbor_of_neighbor
            face_values (o, c, f, W.data ());
            face_values (o, n, nf, Wn.data ());
            for (int q = 0; q < nqf; ++q)
            {
               const int inflow_status = (vel[0] * NORMAL[f][0] + vel[1] * NORMAL[f][1] < 0);      // :125
               cell_shock_ind += inflow_status * (W[q * NC + component] - Wn[q * NC + component]) * face_JxW (o, cl, f, q);
               inflow_measure += inflow_status * face_JxW (o, cl, f, q);
            }
         }
         const double cell_norm = o.cell_average[c * NC + component];            // :189-194
         const double denominator = std::pow (o.diameter (cl), 0.5 * (o.fe.k + 1)) * inflow_measure * cell_norm;
         o.shock_indicator[c] = std::fabs (cell_shock_ind) / denominator;
      }
   }

   //---------------------------------------------------------------------------------------------
   // assemble_explicit.cc:127-248 (+ periodic branch of src_mpi/assemble_explicit.cc:186-260)
   //---------------------------------------------------------------------------------------------
   void integrate_boundary_term (const oracle_ctx &o, int cno, int f, double *local)
   {
      const Cell &cl = o.cells[cno];
      const int D = o.D (), ns = o.fe.ns, nqf = o.face[f].nq;
      std::vector<double> Wplus (nqf * NC), Wminus (nqf * NC), H (nqf * NC);
      face_values (o, cno, f, Wplus.data ());
      double NRM[2], flen;
      o.face_geometry (cl, f, NRM, flen); // fe_v.normal_vector(q): constant along a straight face

      if (cl.nbr[f] >= 0)
      {
         // periodic: neighbour values on its own face, both sides integrate independently
         const int n_cell = cl.nbr[f], n_face = cl.nbr_face[f];
         face_values (o, n_cell, n_face, Wminus.data ());
         for (int q = 0; q < nqf; ++q)
         {
            const int q1 = cl.flip[f] ? nqf - q - 1 : q;
            phys_numerical_flux (o.prm.flux_type, NRM, &Wplus[q * NC], &Wminus[q1 * NC],
                                 &o.cell_average[cno * NC], &o.cell_average[n_cell * NC], &H[q * NC]);
         }
      }
      else
      {
         const int kind = o.prm.bc_kind[cl.bid[f]];
         const double *g = &o.bc_values[(size_t) cl.bface[f] * nqf * NC];
         for (int q = 0; q < nqf; ++q)
         {
            phys_wminus (kind, NRM, &Wplus[q * NC], &g[q * NC], &Wminus[q * NC]);
            const double *Aplus = &o.cell_average[cno * NC];
            double Aminus[NC];
            if (o.prm.compat == ORACLE_COMPAT_MPI) // src_mpi/assemble_explicit.cc:296-321
               phys_wminus (kind, NRM, Aplus, &g[q * NC], Aminus);
            else                                    // src/assemble_explicit.cc:203-204
               for (int c = 0; c < NC; ++c) Aminus[c] = Aplus[c];
            phys_numerical_flux (o.prm.flux_type, NRM, &Wplus[q * NC], &Wminus[q * NC], Aplus,
                                 Aminus, &H[q * NC]);
         }
      }
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns;
         for (int q = 0; q < nqf; ++q)
            F_i += H[q * NC + ci] * o.face[f].phi[(i % ns) * nqf + q] * face_JxW (o, cl, f, q);
         local[i] -= F_i;
      }
   }

   //---------------------------------------------------------------------------------------------
   // assemble_explicit.cc:256-427
   //---------------------------------------------------------------------------------------------
   void integrate_face_term (const oracle_ctx &o, int cno, int f, double *local, double *local_nbr)
   {
      const Cell &cl = o.cells[cno];
      const int ncno = cl.nbr[f], nf = cl.nbr_face[f];
      const Cell &ncl = o.cells[ncno];
      const int D = o.D (), ns = o.fe.ns, nqf = o.face[f].nq;
      std::vector<double> Wplus (nqf * NC), Wminus (nqf * NC), H (nqf * NC);
      face_values (o, cno, f, Wplus.data ());
      face_values (o, ncno, nf, Wminus.data ());
      double NRM[2], flen;
      o.face_geometry (cl, f, NRM, flen);
      // the neighbour's q-th face point is this cell's q-th one unless the two cells run along the line in
      // opposite directions (FEFaceValues of both sides are reinit'ed on the same face: same physical points)
      if (cl.flip[f])
         for (int q = 0; q < nqf / 2; ++q)
            for (int c = 0; c < NC; ++c) std::swap (Wminus[q * NC + c], Wminus[(nqf - 1 - q) * NC + c]);
      for (int q = 0; q < nqf; ++q)
         phys_numerical_flux (o.prm.flux_type, NRM, &Wplus[q * NC], &Wminus[q * NC],
                              &o.cell_average[cno * NC], &o.cell_average[ncno * NC], &H[q * NC]);
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns;
         for (int q = 0; q < nqf; ++q)
            F_i += H[q * NC + ci] * o.face[f].phi[(i % ns) * nqf + q] * face_JxW (o, cl, f, q);
         local[i] -= F_i;
      }
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns;
         for (int q = 0; q < nqf; ++q)
            F_i -= H[q * NC + ci] * o.face[nf].phi[(i % ns) * nqf + (cl.flip[f] ? nqf - q - 1 : q)] * face_JxW (o, ncl, nf, q);
         local_nbr[i] -= F_i;
      }
   }

   // A face with a hanging node, integrated from the FINE side (MeshWorker::loop: "hanging faces integrated from the fine side
   // on sub-faces"): FEFaceValues on the fine cell's face f, FESubfaceValues on child `sub` of the coarse
   // neighbour's face -- the same physical points; assemble_explicit.cc:256-427 does not know the difference.
   void integrate_subface_term (const oracle_ctx &o, int cno, int f, double *local, double *local_nbr)
   {
      const Cell &cl = o.cells[cno];
      const int ncno = cl.nbr[f], nf = cl.nbr_face[f], child = cl.sub[f];
      const int D = o.D (), ns = o.fe.ns, nqf = o.face[f].nq;
      const UnitValues &sv = o.subface[nf][child];
      std::vector<double> Wplus (nqf * NC), Wminus (nqf * NC), H (nqf * NC);
      face_values (o, cno, f, Wplus.data ());
      const double *un = &o.current[(size_t) ncno * D];
      for (int q = 0; q < nqf; ++q)
      {
         const int qn = cl.flip[f] ? nqf - q - 1 : q; // the coarse side's index of this physical point
         for (int c = 0; c < NC; ++c) Wminus[q * NC + c] = 0.0;
         for (int i = 0; i < D; ++i) Wminus[q * NC + i / ns] += un[i] * sv.phi[(i % ns) * nqf + qn];
      }
      double NRM[2], flen;
      o.face_geometry (cl, f, NRM, flen);
      for (int q = 0; q < nqf; ++q)
         phys_numerical_flux (o.prm.flux_type, NRM, &Wplus[q * NC], &Wminus[q * NC], &o.cell_average[cno * NC],
                              &o.cell_average[ncno * NC], &H[q * NC]);
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns;
         for (int q = 0; q < nqf; ++q) F_i += H[q * NC + ci] * o.face[f].phi[(i % ns) * nqf + q] * face_JxW (o, cl, f, q);
         local[i] -= F_i;
      }
      for (int i = 0; i < D; ++i)
      {
         double F_i = 0;
         const int ci = i / ns;
         for (int q = 0; q < nqf; ++q)
            F_i -= H[q * NC + ci] * sv.phi[(i % ns) * nqf + (cl.flip[f] ? nqf - q - 1 : q)] * face_JxW (o, cl, f, q); // JxW of the sub-face = the fine face's
         local_nbr[i] -= F_i;
      }
   }

   // is face f of cell c handled by the boundary worker?  (true boundary or periodic: the two
   // periodic partners share no vertices, cell->at_boundary() stays true)
   bool at_boundary (const oracle_ctx &o, int c, int f, const std::vector<char> &shared)
   {
      (void) o;
      return !shared[4 * c + f];
   }

   //---------------------------------------------------------------------------------------------
   // assemble_explicit.cc:433-452: MeshWorker::loop + ResidualSimple (A7)
   //---------------------------------------------------------------------------------------------
   struct LocalVectors
   {
      // cell vector, then per face an interior and an exterior vector
      std::vector<double> v;
      bool has_int[4], has_ext[4];
   };
}

namespace
{
   void compute_local (const oracle_ctx &o, const std::vector<char> &shared, int c, LocalVectors &L)
   {
      const int D = o.D ();
      L.v.assign ((size_t) 9 * D, 0.0);
      integrate_cell_term (o, c, &L.v[0]);
      for (int f = 0; f < 4; ++f)
      {
         L.has_int[f] = L.has_ext[f] = false;
         if (at_boundary (o, c, f, shared))
         {
            integrate_boundary_term (o, c, f, &L.v[(size_t) (1 + 2 * f) * D]);
            L.has_int[f] = true;
         }
         else if (o.cells[c].hang[f] >= 0)
            ; // the two finer neighbours integrate their halves
         else if (o.cells[c].sub[f] >= 0) // neighbour is coarser: this (fine) side integrates
         {
            integrate_subface_term (o, c, f, &L.v[(size_t) (1 + 2 * f) * D], &L.v[(size_t) (2 + 2 * f) * D]);
            L.has_int[f] = L.has_ext[f] = true;
         }
         else if (o.cells[c].nbr[f] > c) // face integrated once, from the smaller cell
         {
            integrate_face_term (o, c, f, &L.v[(size_t) (1 + 2 * f) * D], &L.v[(size_t) (2 + 2 * f) * D]);
            L.has_int[f] = L.has_ext[f] = true;
         }
      }
   }

   void copy_local (oracle_ctx &o, int c, const LocalVectors &L)
   {
      const int D = o.D ();
      double *r = &o.rhs[(size_t) c * D];
      for (int i = 0; i < D; ++i) r[i] += L.v[i];
      for (int f = 0; f < 4; ++f)
      {
         if (L.has_int[f])
            for (int i = 0; i < D; ++i) r[i] += L.v[(size_t) (1 + 2 * f) * D + i];
         if (L.has_ext[f])
         {
            double *rn = &o.rhs[(size_t) o.cells[c].nbr[f] * D];
            for (int i = 0; i < D; ++i) rn[i] += L.v[(size_t) (2 + 2 * f) * D + i];
         }
      }
   }

   void assemble_system (oracle_ctx &o)
   {
      const std::vector<char> &shared = o.shared;
      std::fill (o.rhs.begin (), o.rhs.end (), 0.0); // right_hand_side = 0, :438
      const int nc = o.cells.size ();
      const int nt = std::max (1, o.prm.n_threads);
      if (nt == 1)
      {
         LocalVectors L;
         for (int c = 0; c < nc; ++c)
         {
            compute_local (o, shared, c, L);
            copy_local (o, c, L);
         }
         return;
      }
      // WorkStream: workers in parallel on a chunk, copier serial and in mesh order
      const int chunk = 64 * nt;
      std::vector<LocalVectors> scratch (chunk);
      for (int c0 = 0; c0 < nc; c0 += chunk)
      {
         const int c1 = std::min (nc, c0 + chunk);
#pragma omp parallel for num_threads(nt) schedule(static)
         for (int c = c0; c < c1; ++c) compute_local (o, shared, c, scratch[c - c0]);
         for (int c = c0; c < c1; ++c) copy_local (o, c, scratch[c - c0]);
      }
   }

   //---------------------------------------------------------------------------------------------
   // claw.cc:562-597
   //---------------------------------------------------------------------------------------------
   void compute_cell_average (oracle_ctx &o)
   {
      const int D = o.D (), ns = o.fe.ns, nq = o.vol.nq;
      std::vector<double> val (nq * NC);
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         const Cell &cl = o.cells[c];
         const double *u = &o.current[c * D];
         // get_function_values
         for (int q = 0; q < nq; ++q)
         {
            for (int k = 0; k < NC; ++k) val[q * NC + k] = 0.0;
            for (int i = 0; i < D; ++i) val[q * NC + i / ns] += u[i] * o.vol.phi[(i % ns) * nq + q];
         }
         double *avg = &o.cell_average[c * NC];
         for (int k = 0; k < NC; ++k) avg[k] = 0.0;
         for (int q = 0; q < nq; ++q)
            for (int k = 0; k < NC; ++k) avg[k] += val[q * NC + k] * o.JxW (cl, q);
         const double measure = o.measure (cl);
         for (int k = 0; k < NC; ++k) avg[k] /= measure;
      }
   }

   //---------------------------------------------------------------------------------------------
   // claw.cc:484-511 and 444-478
   //---------------------------------------------------------------------------------------------
   double compute_time_step (oracle_ctx &o, double elapsed, double final_time, double time_step)
   {
      o.global_dt = 1.0e20;
      if (o.q1 ())
      {
         // compute_time_step_q, claw.cc:518-557: largest eigenvalue over the 4x4 equispaced points of QIterated(QTrapez,3)
         const int D = o.D (), ns = o.fe.ns, nq = o.dtq.nq;
         for (size_t c = 0; c < o.cells.size (); ++c)
         {
            const double *u = &o.current[c * D];
            double max_eigenvalue = 0.0;
            for (int q = 0; q < nq; ++q)
            {
               double W[NC] = {0, 0, 0, 0};
               for (int i = 0; i < D; ++i) W[i / ns] += u[i] * o.dtq.phi[(i % ns) * nq + q];
               max_eigenvalue = std::max (max_eigenvalue, phys_max_eigenvalue (W));
            }
            const double h = o.diameter (o.cells[c]) / std::sqrt (2.0);
            o.dt[c] = o.prm.cfl * h / max_eigenvalue / (2.0 * o.fe.k + 1.0);
            o.global_dt = std::min (o.global_dt, o.dt[c]);
         }
      }
      else
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         const double h = o.diameter (o.cells[c]) / std::sqrt (2.0);
         const double *avg = &o.cell_average[c * NC];
         const double sonic = phys_sound_speed (avg);
         const double density = avg[RHO];
         double max_eigenvalue = 0.0;
         for (int d = 0; d < 2; ++d) max_eigenvalue += (sonic + std::fabs (avg[d] / density)) / h;
         o.dt[c] = o.prm.cfl / max_eigenvalue / (2.0 * o.fe.k + 1.0);
         o.global_dt = std::min (o.global_dt, o.dt[c]);
      }
      if (o.prm.local_time_step) return o.global_dt; // claw.cc:469: what follows is for "global" only; dt(c) stays per cell
      if (o.global_dt > 0 && time_step > 0) o.global_dt = std::min (o.global_dt, time_step);
      if (elapsed + o.global_dt > final_time) o.global_dt = final_time - elapsed;
      std::fill (o.dt.begin (), o.dt.end (), o.global_dt);
      return o.global_dt;
   }

   //---------------------------------------------------------------------------------------------
   // limiter.cc:15-30
   //---------------------------------------------------------------------------------------------
   double minmod (const double &a, const double &b, const double &c, const double &Mdx2)
   {
      const double aa = std::fabs (a);
      if (aa < Mdx2) return a;
      if (a * b > 0 && b * c > 0)
      {
         const double s = (a > 0) ? 1.0 : -1.0;
         return s * std::min (aa, std::min (std::fabs (b), std::fabs (c)));
      }
      return 0;
   }

   // Common part of limiter.cc:283-345 / 425-486: differences of averages, characteristic
   // projection, minmod.  Returns true if the limiter is active.
   struct TVBWork
   {
      double Dx[NC], Dy[NC], Dx_new[NC], Dy_new[NC];
      double Rx[16], Lx[16], Ry[16], Ly[16];
   };

   bool tvb_slopes (const oracle_ctx &o, int c, double beta, TVBWork &w)
   {
      const Cell &cl = o.cells[c];
      const double dx = o.diameter (cl) / std::sqrt (2.0);
      const double Mdx2 = o.prm.M * dx * dx;
      const double *avg = &o.cell_average[c * NC];
      double dbx[NC], dfx[NC], dby[NC], dfy[NC];
      // lcell/rcell/bcell/tcell (claw.cc:336-380) = neighbours through faces 0/1/2/3; periodic
      // partners count as neighbours too (the only tree with periodic faces resolves them
      // through the periodic map, src_mpi/claw.cc:417-465)
      const bool has[4] = {cl.nbr[0] >= 0, cl.nbr[1] >= 0, cl.nbr[2] >= 0, cl.nbr[3] >= 0};
      for (int i = 0; i < NC; ++i)
      {
         dbx[i] = w.Dx[i];
         dfx[i] = w.Dx[i];
         dby[i] = w.Dy[i];
         dfy[i] = w.Dy[i];
      }
      if (has[0])
         for (int i = 0; i < NC; ++i) dbx[i] = avg[i] - o.cell_average[cl.nbr[0] * NC + i];
      if (has[1])
         for (int i = 0; i < NC; ++i) dfx[i] = o.cell_average[cl.nbr[1] * NC + i] - avg[i];
      if (has[2])
         for (int i = 0; i < NC; ++i) dby[i] = avg[i] - o.cell_average[cl.nbr[2] * NC + i];
      if (has[3])
         for (int i = 0; i < NC; ++i) dfy[i] = o.cell_average[cl.nbr[3] * NC + i] - avg[i];

      if (o.prm.char_lim)
      {
         phys_eigen (avg, w.Rx, w.Lx, w.Ry, w.Ly);
         phys_to_char (w.Lx, dbx);
         phys_to_char (w.Lx, dfx);
         phys_to_char (w.Ly, dby);
         phys_to_char (w.Ly, dfy);
         phys_to_char (w.Lx, w.Dx);
         phys_to_char (w.Ly, w.Dy);
      }
      double change_x = 0, change_y = 0;
      for (int i = 0; i < NC; ++i)
      {
         w.Dx_new[i] = minmod (w.Dx[i], beta * dbx[i], beta * dfx[i], Mdx2);
         w.Dy_new[i] = minmod (w.Dy[i], beta * dby[i], beta * dfy[i], Mdx2);
         change_x += std::fabs (w.Dx_new[i] - w.Dx[i]);
         change_y += std::fabs (w.Dy_new[i] - w.Dy[i]);
      }
      change_x /= NC;
      change_y /= NC;
      return change_x + change_y > 1.0e-10;
   }

   //---------------------------------------------------------------------------------------------
   // limiter.cc:224-370
   //---------------------------------------------------------------------------------------------
   void apply_limiter_TVB_Qk (oracle_ctx &o)
   {
      if (o.fe.k == 0) return;
      const int D = o.D (), ns = o.fe.ns, nq = o.vol.nq;
      std::vector<double> grad (nq * NC * 2);
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         if (!(o.shock_indicator[c] > 1.0)) continue;
         const Cell &cl = o.cells[c];
         double *u = &o.current[c * D];
         const double dx = o.diameter (cl) / std::sqrt (2.0);
         const double h[2] = {cl.hx, cl.hy};
         // get_function_gradients at QGauss(k+1)
         for (int q = 0; q < nq; ++q)
         {
            for (int k = 0; k < NC * 2; ++k) grad[q * NC * 2 + k] = 0.0;
            for (int i = 0; i < D; ++i)
               for (int d = 0; d < 2; ++d)
                  grad[(q * NC + i / ns) * 2 + d] += u[i] * (o.vol.dphi[((i % ns) * nq + q) * 2 + d] / h[d]);
         }
         TVBWork w;
         for (int i = 0; i < NC; ++i)
         {
            double avg_grad[2] = {0, 0};
            for (int q = 0; q < nq; ++q)
               for (int d = 0; d < 2; ++d) avg_grad[d] += grad[(q * NC + i) * 2 + d] * o.JxW (cl, q);
            for (int d = 0; d < 2; ++d) avg_grad[d] /= (cl.hx * cl.hy);
            w.Dx[i] = dx * avg_grad[0];
            w.Dy[i] = dx * avg_grad[1];
         }
         if (tvb_slopes (o, c, o.prm.beta, w))
         {
            for (int i = 0; i < NC; ++i)
            {
               w.Dx_new[i] /= dx;
               w.Dy_new[i] /= dx;
            }
            if (o.prm.char_lim)
            {
               phys_to_con (w.Rx, w.Dx_new);
               phys_to_con (w.Ry, w.Dy_new);
            }
            const double xc = cl.x0 + 0.5 * cl.hx, yc = cl.y0 + 0.5 * cl.hy;
            for (int i = 0; i < D; ++i)
            {
               const int comp_i = i / ns;
               const double px = cl.x0 + o.support.x[i % ns] * cl.hx, py = cl.y0 + o.support.y[i % ns] * cl.hy;
               const double dr[2] = {px - xc, py - yc};
               u[i] = o.cell_average[c * NC + comp_i] + dr[0] * w.Dx_new[comp_i] + dr[1] * w.Dy_new[comp_i];
            }
            o.limited[c] |= 1;
         }
      }
   }

   //---------------------------------------------------------------------------------------------
   // limiter.cc:376-516
   //---------------------------------------------------------------------------------------------
   void apply_limiter_TVB_Pk (oracle_ctx &o)
   {
      if (o.fe.k == 0) return;
      const int D = o.D (), ns = o.fe.ns, k = o.fe.k;
      static const double sqrt_3 = std::sqrt (3.0);
      const double beta = 0.5 * o.prm.beta;
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         if (!(o.shock_indicator[c] > 1.0)) continue;
         double *u = &o.current[c * D];
         TVBWork w;
         for (int i = 0; i < D; ++i)
         {
            const int comp_i = i / ns, base_i = i % ns;
            if (base_i == 1)
               w.Dx[comp_i] = u[i] * sqrt_3;
            else if (base_i == k + 1)
               w.Dy[comp_i] = u[i] * sqrt_3;
         }
         const double ang_mom = w.Dx[1] - w.Dy[0];
         if (tvb_slopes (o, c, beta, w))
         {
            if (o.prm.char_lim)
            {
               phys_to_con (w.Rx, w.Dx_new);
               phys_to_con (w.Ry, w.Dy_new);
            }
            if (o.prm.conserve_angular_momentum)
            {
               w.Dy_new[0] = 0.5 * (w.Dy_new[0] - (ang_mom - w.Dx_new[1]));
               w.Dx_new[1] = ang_mom + w.Dy_new[0];
            }
            for (int i = 0; i < D; ++i)
            {
               const int comp_i = i / ns, base_i = i % ns;
               if (base_i == 1)
                  u[i] = w.Dx_new[comp_i] / sqrt_3;
               else if (base_i == k + 1)
                  u[i] = w.Dy_new[comp_i] / sqrt_3;
               else if (base_i != 0)
                  u[i] = 0.0;
            }
            o.limited[c] |= 1;
         }
      }
   }

   //---------------------------------------------------------------------------------------------
   // src_mpi/limiter.cc:400-553, the Barth-Jespersen type "minmax" limiter of the MPI tree (Qk only,
   // src_mpi/parameters.cc:610-611).  Restated as written, including two properties a reader might
   // not expect: (1) without characteristic limiting avg_min / avg_max start from the
   // zero-initialised Vector<double>(n_components) (limiter.cc:438, 446-451 assign them only under
   // char_lim), so the bounds always contain 0; (2) only faces with !cell->at_boundary() contribute
   // neighbours (:453-454), i.e. periodic partners do not.  The mean gradient uses QGauss(nq) with
   // nq = k/2+1 (k even) or (k+1)/2+1 (k odd) (:407-412).
   //---------------------------------------------------------------------------------------------
   void apply_limiter_minmax_Qk (oracle_ctx &o)
   {
      if (o.fe.k == 0) return;
      const int D = o.D (), ns = o.fe.ns, nq = o.mmgrad.nq;
      std::vector<double> grad (nq * NC * 2);
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         if (!(o.shock_indicator[c] > 1.0)) continue;                               // :437
         const Cell &cl = o.cells[c];
         double *u = &o.current[c * D];
         const double dx = o.diameter (cl) / std::sqrt (2.0);
         const double Mdx2 = o.prm.M * dx * dx;
         const double h[2] = {cl.hx, cl.hy};
         double avg_min[NC] = {0, 0, 0, 0}, avg_max[NC] = {0, 0, 0, 0}, avg_cell[NC], avg_nbr[NC];
         for (int i = 0; i < NC; ++i) avg_cell[i] = o.cell_average[c * NC + i];
         double Rx[16], Lx[16];
         if (o.prm.char_lim)                                                        // :446-452
         {
            // compute_eigen_matrix (W, R, L): the streamline-direction matrices, src_mpi/equation.h:299-335
            phys_eigen_stream_restated (&o.cell_average[c * NC], Rx, Lx);
            phys_to_char (Lx, avg_cell);
            for (int i = 0; i < NC; ++i) avg_min[i] = avg_max[i] = avg_cell[i];
         }
         for (int f = 0; f < 4; ++f)
            if (o.shared[c * 4 + f])                                                // :454
            {
               for (int i = 0; i < NC; ++i) avg_nbr[i] = o.cell_average[cl.nbr[f] * NC + i];
               if (o.prm.char_lim) phys_to_char (Lx, avg_nbr);
               for (int i = 0; i < NC; ++i)
               {
                  avg_min[i] = std::min (avg_min[i], avg_nbr[i]);
                  avg_max[i] = std::max (avg_max[i], avg_nbr[i]);
               }
            }
         // get_function_gradients at QGauss(nq), :473-490
         for (int q = 0; q < nq; ++q)
         {
            for (int k = 0; k < NC * 2; ++k) grad[q * NC * 2 + k] = 0.0;
            for (int i = 0; i < D; ++i)
               for (int d = 0; d < 2; ++d)
                  grad[(q * NC + i / ns) * 2 + d] += u[i] * (o.mmgrad.dphi[((i % ns) * nq + q) * 2 + d] / h[d]);
         }
         double dumin[NC], dumax[NC], Dx[NC], Dy[NC];
         for (int i = 0; i < NC; ++i)
         {
            dumin[i] = avg_min[i] - avg_cell[i];
            dumax[i] = avg_max[i] - avg_cell[i];
            double avg_grad[2] = {0, 0};
            for (int q = 0; q < nq; ++q)
               for (int d = 0; d < 2; ++d) avg_grad[d] += grad[(q * NC + i) * 2 + d] * (o.mmgrad.w[q] * cl.hx * cl.hy);
            for (int d = 0; d < 2; ++d) avg_grad[d] /= (cl.hx * cl.hy);
            Dx[i] = avg_grad[0];
            Dy[i] = avg_grad[1];
         }
         if (o.prm.char_lim)                                                        // :491-495
         {
            phys_to_char (Lx, Dx);
            phys_to_char (Lx, Dy);
         }
         double theta[NC] = {1.0, 1.0, 1.0, 1.0};
         const double xc = cl.x0 + 0.5 * cl.hx, yc = cl.y0 + 0.5 * cl.hy;
         for (int f = 0; f < 4; ++f)                                                // :498-510, all four faces
         {
            const double fx = (f == 0) ? cl.x0 : (f == 1) ? cl.x0 + cl.hx : xc;
            const double fy = (f == 2) ? cl.y0 : (f == 3) ? cl.y0 + cl.hy : yc;
            const double dr[2] = {fx - xc, fy - yc};
            for (int i = 0; i < NC; ++i)
               if (dumax[i] - dumin[i] > Mdx2)
               {
                  const double du = dr[0] * Dx[i] + dr[1] * Dy[i];
                  if (du > 0.0)
                     theta[i] = std::min (theta[i], dumax[i] / du);
                  else if (du < 0.0)
                     theta[i] = std::min (theta[i], dumin[i] / du);
               }
         }
         double change = 0;
         for (int i = 0; i < NC; ++i) change += theta[i];
         change /= NC;
         if (change < 0.99)                                                         // :519
         {
            for (int i = 0; i < NC; ++i)
            {
               Dx[i] *= theta[i];
               Dy[i] *= theta[i];
            }
            if (o.prm.char_lim)
            {
               phys_to_con (Rx, Dx);
               phys_to_con (Rx, Dy);
            }
            for (int i = 0; i < D; ++i)
            {
               const int comp_i = i / ns;
               const double px = cl.x0 + o.support.x[i % ns] * cl.hx, py = cl.y0 + o.support.y[i % ns] * cl.hy;
               const double dr[2] = {px - xc, py - yc};
               u[i] = o.cell_average[c * NC + comp_i] + dr[0] * Dx[comp_i] + dr[1] * Dy[comp_i];
            }
            o.limited[c] |= 1;
         }
      }
   }

   void apply_limiter (oracle_ctx &o) // limiter.cc:35-65, src_mpi/limiter.cc:36-70
   {
      if (o.prm.limiter_type == ORACLE_LIMITER_MINMAX)
      {
         if (o.prm.basis == ORACLE_BASIS_QK) apply_limiter_minmax_Qk (o);
         return;
      }
      if (o.prm.limiter_type != ORACLE_LIMITER_TVB) return;
      if (o.prm.basis == ORACLE_BASIS_QK)
         apply_limiter_TVB_Qk (o);
      else
         apply_limiter_TVB_Pk (o);
   }

   //---------------------------------------------------------------------------------------------
   // positivity.cc:16-208
   //---------------------------------------------------------------------------------------------
   void point_values (const oracle_ctx &o, const UnitValues &uv, const double *u, int comp, double *out)
   {
      const int ns = o.fe.ns;
      for (int q = 0; q < uv.nq; ++q)
      {
         double v = 0.0;
         for (int m = 0; m < ns; ++m) v += u[comp * ns + m] * uv.phi[m * uv.nq + q];
         out[q] = v;
      }
   }

   int apply_positivity_limiter (oracle_ctx &o)
   {
      if (o.fe.k == 0) return 0;
      const double gas_gamma = 1.4;
      const double eps = 1.0e-13;
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         double eps1 = o.cell_average[c * NC + RHO];
         const double pressure = phys_pressure (&o.cell_average[c * NC]);
         eps1 = std::min (eps1, pressure);
         if (eps1 < eps) return -1; // "Fatal: Negative states", positivity.cc:33-37
      }
      const int D = o.D (), ns = o.fe.ns, nqp = o.posx.nq;
      std::vector<double> density_values (nqp), energy_values (nqp), mx (nqp), my (nqp);
      for (size_t c = 0; c < o.cells.size (); ++c)
      {
         double *u = &o.current[c * D];
         double rho_min = 1.0e20;
         point_values (o, o.posx, u, RHO, density_values.data ());
         for (int q = 0; q < nqp; ++q) rho_min = std::min (rho_min, density_values[q]);
         point_values (o, o.posy, u, RHO, density_values.data ());
         for (int q = 0; q < nqp; ++q) rho_min = std::min (rho_min, density_values[q]);

         const double density_average = o.cell_average[c * NC + RHO];
         const double rat = std::fabs (density_average - eps) / (std::fabs (density_average - rho_min) + 1.0e-13);
         const double theta1 = std::min (rat, 1.0);
         if (theta1 < 1.0)
         {
            o.limited[c] |= 2;
            if (o.prm.basis == ORACLE_BASIS_QK)
            {
               for (int i = 0; i < D; ++i)
                  if (i / ns == RHO) u[i] = theta1 * u[i] + (1.0 - theta1) * density_average;
            }
            else
            {
               for (int i = 0; i < D; ++i)
                  if (i / ns == RHO && i % ns > 0) u[i] *= theta1;
            }
         }

         const double energy_average = o.cell_average[c * NC + ENE];
         const double momentum_average[2] = {o.cell_average[c * NC + 0], o.cell_average[c * NC + 1]};
         double theta2 = 1.0;
         for (int d = 0; d < 2; ++d)
         {
            const UnitValues &uv = (d == 0) ? o.posx : o.posy;
            point_values (o, uv, u, RHO, density_values.data ());
            point_values (o, uv, u, 0, mx.data ());
            point_values (o, uv, u, 1, my.data ());
            point_values (o, uv, u, ENE, energy_values.data ());
            for (int q = 0; q < nqp; ++q)
            {
               const double nrm2 = mx[q] * mx[q] + my[q] * my[q];
               const double pressure = (gas_gamma - 1.0) * (energy_values[q] - 0.5 * nrm2 / density_values[q]);
               if (pressure < eps)
               {
                  const double drho = density_values[q] - density_average;
                  const double dm[2] = {mx[q] - momentum_average[0], my[q] - momentum_average[1]};
                  const double dE = energy_values[q] - energy_average;
                  const double a1 = 2.0 * drho * dE - (dm[0] * dm[0] + dm[1] * dm[1]);
                  double b1 = 2.0 * drho * (energy_average - eps / (gas_gamma - 1.0)) + 2.0 * density_average * dE
                              - 2.0 * (momentum_average[0] * dm[0] + momentum_average[1] * dm[1]);
                  double c1 = 2.0 * density_average * energy_average
                              - (momentum_average[0] * momentum_average[0] + momentum_average[1] * momentum_average[1])
                              - 2.0 * eps * density_average / (gas_gamma - 1.0);
                  b1 /= a1;
                  c1 /= a1;
                  const double Dd = std::sqrt (std::fabs (b1 * b1 - 4.0 * c1));
                  const double t1 = 0.5 * (-b1 - Dd);
                  const double t2 = 0.5 * (-b1 + Dd);
                  double t;
                  if (t1 > -1.0e-12 && t1 < 1.0 + 1.0e-12)
                     t = t1;
                  else if (t2 > -1.0e-12 && t2 < 1.0 + 1.0e-12)
                     t = t2;
                  else
                     return -2; // "Problem in positivity limiter", positivity.cc:160-169
                  t = std::min (1.0, t);
                  t = std::max (0.0, t);
                  if (std::fabs (1.0 - t) < 1.0e-14) t = 0.0;
                  theta2 = std::min (theta2, t);
               }
            }
         }
         if (theta2 < 1.0)
         {
            o.limited[c] |= 4;
            if (o.prm.basis == ORACLE_BASIS_QK)
            {
               for (int i = 0; i < D; ++i) u[i] = theta2 * u[i] + (1.0 - theta2) * o.cell_average[c * NC + i / ns];
            }
            else
            {
               for (int i = 0; i < D; ++i)
                  if (i % ns > 0) u[i] *= theta2;
            }
         }
      }
      return 0;
   }

   //---------------------------------------------------------------------------------------------
   // claw.cc:747-766 (one pass of the rk loop)
   //---------------------------------------------------------------------------------------------
   int rk_stage (oracle_ctx &o, int rk, double dt, double *res_norm)
   {
      assemble_system (o);
      if (res_norm)
      {
         double s = 0;
         for (double r : o.rhs) s += r * r;
         *res_norm = std::sqrt (s);
      }
      const int D = o.D ();
      // solve(), rk3 branch: claw.cc:694-713
      for (size_t c = 0; c < o.cells.size (); ++c)
         for (int i = 0; i < D; ++i)
            o.newton_update[c * D + i] = (o.prm.local_time_step ? o.dt[c] : dt) * o.rhs[c * D + i] * o.inv_mass[c * D + i]; // dt(cell_no), :709
      const double a = o.ark[rk];
      for (size_t j = 0; j < o.current.size (); ++j) o.current[j] += o.newton_update[j];           // :757
      for (size_t j = 0; j < o.current.size (); ++j) o.current[j] = (1.0 - a) * o.current[j] + a * o.old[j]; // :760
      compute_cell_average (o);                                                                     // :762
      compute_shock_indicator (o);                                                                  // :763
      std::fill (o.limited.begin (), o.limited.end (), 0);
      apply_limiter (o);                                                                            // :764
      if (o.prm.pos_lim) return apply_positivity_limiter (o);                                       // :766
      return 0;
   }
}

//-----------------------------------------------------------------------------------------------
// C interface
//-----------------------------------------------------------------------------------------------
extern "C" {

const char *oracle_last_error (void) { return g_error.c_str (); }

oracle_ctx *oracle_create (int nv, const double *V, int nc, const int *C, int nb, const int *BL,
                           const int *BID, const oracle_params *prm)
{
   oracle_ctx *o = new oracle_ctx;
   o->prm = *prm;
   if (!build_mesh (*o, nv, V, nc, C, nb, BL, BID))
   {
      delete o;
      return nullptr;
   }
   FE &fe = o->fe;
   fe.init (prm->basis, prm->degree);
   const int n1 = fe.n1;
   // QGauss<2>(k+1): tensor product, x fastest
   {
      std::vector<double> x, y, w;
      for (int b = 0; b < n1; ++b)
         for (int a = 0; a < n1; ++a)
         {
            x.push_back (fe.gx[a]);
            y.push_back (fe.gx[b]);
            w.push_back (fe.gw[a] * fe.gw[b]);
         }
      o->vol.init (fe, x, y, w);
      o->support.init (fe, x, y, w); // Qk: support points == Gauss points, same order (A3)
   }
   {
      const int k = fe.k;
      const int nq = (k % 2 == 0) ? k / 2 + 1 : (k + 1) / 2 + 1; // src_mpi/limiter.cc:407-412
      std::vector<double> gx, gw, x, y, w;
      gauss_rule (nq, gx, gw);
      for (int b = 0; b < nq; ++b)
         for (int a = 0; a < nq; ++a)
         {
            x.push_back (gx[a]);
            y.push_back (gx[b]);
            w.push_back (gw[a] * gw[b]);
         }
      o->mmgrad.init (fe, x, y, w);
   }
   for (int f = 0; f < 4; ++f)
   {
      std::vector<double> x, y, w;
      for (int q = 0; q < n1; ++q)
      {
         const double t = fe.gx[q];
         x.push_back (f == 0 ? 0.0 : f == 1 ? 1.0 : t);
         y.push_back (f == 2 ? 0.0 : f == 3 ? 1.0 : t);
         w.push_back (fe.gw[q]);
      }
      o->face[f].init (fe, x, y, w);
   }
   {
      const int k = fe.k;
      const int N = (k + 3) % 2 == 0 ? (k + 3) / 2 : (k + 4) / 2; // positivity.cc:43
      std::vector<double> gl;
      gauss_lobatto_rule (N, gl);
      std::vector<double> x, y, w;
      for (int b = 0; b < n1; ++b) // Quadrature<2>(QGaussLobatto(N), QGauss(k+1)): first along x
         for (int a = 0; a < N; ++a)
         {
            x.push_back (gl[a]);
            y.push_back (fe.gx[b]);
            w.push_back (0.0);
         }
      o->posx.init (fe, x, y, w);
      x.clear ();
      y.clear ();
      for (int b = 0; b < N; ++b)
         for (int a = 0; a < n1; ++a)
         {
            x.push_back (fe.gx[a]);
            y.push_back (gl[b]);
         }
      o->posy.init (fe, x, y, w);
   }
   for (int f = 0; f < 4; ++f)
      for (int child = 0; child < 2; ++child)
      {
         std::vector<double> x, y, w;
         for (int q = 0; q < n1; ++q)
         {
            const double t = 0.5 * (fe.gx[q] + child);
            x.push_back (f == 0 ? 0.0 : f == 1 ? 1.0 : t);
            y.push_back (f == 2 ? 0.0 : f == 3 ? 1.0 : t);
            w.push_back (0.5 * fe.gw[q]);
         }
         o->subface[f][child].init (fe, x, y, w);
      }
   if (!o->hanging.empty () && (prm->limiter_type != ORACLE_LIMITER_NONE || prm->shock_indicator != 0))
   {
      // the positivity limiter is local to a cell and runs next to hanging nodes as it is
      g_error = "hanging nodes: no TVB / minmax limiter or shock indicator (their neighbour lists are same-level)";
      delete o;
      return nullptr;
   }
   {
      // QIterated<2>(QTrapez<1>(), 3): points i/3, tensor product, x fastest (claw.cc:522)
      std::vector<double> x, y, w;
      for (int b = 0; b < 4; ++b)
         for (int a = 0; a < 4; ++a)
         {
            x.push_back (a / 3.0);
            y.push_back (b / 3.0);
            w.push_back (0.0);
         }
      o->dtq.init (fe, x, y, w);
   }
   if (o->q1 () && (prm->basis != ORACLE_BASIS_QK || prm->limiter_type != ORACLE_LIMITER_NONE))
   {
      // parameters.cc:545-549: TVB and Pk need Cartesian grids.  The positivity limiter (positivity.cc) works on unit-cell
      // point values and the mapped cell average, so it needs nothing beyond what apply_positivity_limiter does
      g_error = "mapping = q1: Qk basis without the TVB limiter only";
      delete o;
      return nullptr;
   }
   const size_t ndof = (size_t) nc * fe.D;
   o->current.assign (ndof, 0.0);
   o->old.assign (ndof, 0.0);
   o->rhs.assign (ndof, 0.0);
   o->newton_update.assign (ndof, 0.0);
   o->dt.assign (nc, 0.0);
   o->cell_average.assign ((size_t) nc * NC, 0.0);
   o->shock_indicator.assign (nc, 1e20);
   o->limited.assign (nc, 0);
   o->bc_values.assign ((size_t) std::max (1, o->n_bfaces) * n1 * NC, 0.0);
   o->global_dt = 0;
   // claw.cc:141-159
   if (fe.k == 0) { o->ark[0] = 0.0; o->n_rk = 1; }
   else if (fe.k == 1) { o->ark[0] = 0.0; o->ark[1] = 1.0 / 2.0; o->n_rk = 2; }
   else { o->ark[0] = 0.0; o->ark[1] = 3.0 / 4.0; o->ark[2] = 1.0 / 3.0; o->n_rk = 3; }
   compute_inv_mass_matrix (*o);
   return o;
}

void oracle_destroy (oracle_ctx *o)
{
   delete o;
}

int oracle_n_cells (const oracle_ctx *o) { return o->cells.size (); }
int oracle_dofs_per_cell (const oracle_ctx *o) { return o->fe.D; }
int oracle_n_q_face (const oracle_ctx *o) { return o->fe.n1; }
int oracle_n_q_cell (const oracle_ctx *o) { return o->vol.nq; }
int oracle_n_bfaces (const oracle_ctx *o) { return o->n_bfaces; }
int oracle_n_rk (const oracle_ctx *o) { return o->n_rk; }
double oracle_ark (const oracle_ctx *o, int rk) { return o->ark[rk]; }

void oracle_get_neighbors (const oracle_ctx *o, int *nbr)
{
   for (size_t c = 0; c < o->cells.size (); ++c)
      for (int f = 0; f < 4; ++f)
         nbr[4 * c + f] = o->cells[c].nbr[f] >= 0 ? o->cells[c].nbr[f] : -1 - o->cells[c].bface[f];
}

void oracle_get_bfaces (const oracle_ctx *o, int *cell, int *face, int *bid, double *xq)
{
   const int nqf = o->fe.n1;
   for (int b = 0; b < o->n_bfaces; ++b)
   {
      const int c = o->bf_cell[b], f = o->bf_face[b];
      cell[b] = c;
      face[b] = f;
      bid[b] = o->bf_bid[b];
      const Cell &cl = o->cells[c];
      for (int q = 0; q < nqf; ++q)
         o->map_point (cl, o->face[f].x[q], o->face[f].y[q], xq[(b * nqf + q) * 2 + 0], xq[(b * nqf + q) * 2 + 1]);
   }
}

void oracle_get_cell_qpoints (const oracle_ctx *o, double *xq)
{
   const int nq = o->vol.nq;
   for (size_t c = 0; c < o->cells.size (); ++c)
      for (int q = 0; q < nq; ++q)
         o->map_point (o->cells[c], o->vol.x[q], o->vol.y[q], xq[(c * nq + q) * 2 + 0], xq[(c * nq + q) * 2 + 1]);
}

void oracle_get_tables (const oracle_ctx *o, double *gx, double *gw)
{
   for (int a = 0; a < o->fe.n1; ++a)
   {
      gx[a] = o->fe.gx[a];
      gw[a] = o->fe.gw[a];
   }
}

void oracle_set_initial_condition (oracle_ctx *o, const double *f)
{
   const int D = o->D (), ns = o->fe.ns, nq = o->vol.nq;
   for (size_t c = 0; c < o->cells.size (); ++c)
   {
      double *u = &o->old[c * D];
      if (o->prm.basis == ORACLE_BASIS_QK)
      {
         // VectorTools::interpolate: value at the support point of DoF i (ic.cc:104-120)
         for (int i = 0; i < D; ++i) u[i] = f[(c * nq + i % ns) * NC + i / ns];
      }
      else
      {
         // create_right_hand_side then divide by the diagonal mass (ic.cc:128-168)
         for (int i = 0; i < D; ++i)
         {
            double r = 0.0;
            for (int q = 0; q < nq; ++q)
               r += f[(c * nq + q) * NC + i / ns] * o->vol.phi[(i % ns) * nq + q] * o->JxW (o->cells[c], q);
            u[i] = r * o->inv_mass[c * D + i];
         }
      }
   }
   o->current = o->old;
}

void oracle_set_solution (oracle_ctx *o, const double *u)
{
   std::copy (u, u + o->current.size (), o->current.begin ());
   o->old = o->current;
}

void oracle_get_solution (const oracle_ctx *o, double *u) { std::copy (o->current.begin (), o->current.end (), u); }
void oracle_commit_step (oracle_ctx *o) { o->old = o->current; }
/* external force values at the cell quadrature points (parameters.external_force.vector_value_list,
 * src_mpi/assemble_explicit.cc:56-58); NULL restores the hard-wired forcing of src/ */
void oracle_set_external_force (oracle_ctx *o, const double *f)
{
   if (!f)
      o->ext_force.clear ();
   else
      o->ext_force.assign (f, f + (size_t) o->cells.size () * o->vol.nq * 2);
}
void oracle_set_bc_values (oracle_ctx *o, const double *g)
{
   std::copy (g, g + (size_t) o->n_bfaces * o->fe.n1 * NC, o->bc_values.begin ());
}

void oracle_compute_cell_average (oracle_ctx *o) { compute_cell_average (*o); }
void oracle_get_cell_average (const oracle_ctx *o, double *avg)
{
   std::copy (o->cell_average.begin (), o->cell_average.end (), avg);
}
void oracle_assemble (oracle_ctx *o) { assemble_system (*o); }
void oracle_get_rhs (const oracle_ctx *o, double *rhs) { std::copy (o->rhs.begin (), o->rhs.end (), rhs); }
double oracle_compute_dt (oracle_ctx *o, double elapsed, double final_time, double time_step)
{
   return compute_time_step (*o, elapsed, final_time, time_step);
}
void oracle_compute_shock_indicator (oracle_ctx *o) { compute_shock_indicator (*o); }
void oracle_get_shock_indicator (const oracle_ctx *o, double *ind) { std::copy (o->shock_indicator.begin (), o->shock_indicator.end (), ind); }
void oracle_apply_limiter (oracle_ctx *o)
{
   compute_shock_indicator (*o); // claw.cc:1000-1001
   std::fill (o->limited.begin (), o->limited.end (), 0);
   apply_limiter (*o);
}
int oracle_apply_positivity (oracle_ctx *o) { return apply_positivity_limiter (*o); }
void oracle_get_limited_flags (const oracle_ctx *o, int *flags) { std::copy (o->limited.begin (), o->limited.end (), flags); }
int oracle_rk_stage (oracle_ctx *o, int rk, double dt, double *res_norm) { return rk_stage (*o, rk, dt, res_norm); }

double oracle_run_steps (oracle_ctx *o, int n_steps, int *err)
{
   double t = 0.0;
   if (err) *err = 0;
   for (int s = 0; s < n_steps; ++s)
   {
      const double dt = compute_time_step (*o, t, 1.0e20, -1.0);
      for (int rk = 0; rk < o->n_rk; ++rk)
      {
         const int e = rk_stage (*o, rk, dt, nullptr);
         if (e && err) { *err = e; return t; }
      }
      t += dt;
      o->old = o->current;
   }
   return t;
}
}
