// CPU check of adapters/dealii_flatten.h against adapters/dealii_mock (a MOCK of deal.II, see there): the adapter must
// produce, from "deal.II" iterators, the flat mesh that the library's own host code derives for the same rectangle
// (dflo_mesh_create "rectangle" + dflo_mesh_flatten), whatever order deal.II numbers the cells in; the dof map and the
// captured boundary expressions are checked, and the resulting structures are handed to dflo_b200_create (on a box
// without a GPU that call must fail with DFLO_E_NO_DEVICE -- the adapter has no CPU path of its own either).
//   g++ -std=c++17 -Iinclude -Iadapters -Iadapters/dealii_mock adapters/test_adapter.cc -Ldflo_b200/csrc -ldflo_b200
#include <deal.II/mock.h>

#include "dealii_flatten.h"
#include <dflo_host.h>

#include <algorithm>
#include <cstdio>
#include <numeric>

// the members of Parameters::AllParameters<2> the adapter reads (src/parameters.h:227-256, 364-407)
struct MockAllParameters
{
   enum FluxType { lxf, sw, kfvs, roe, hllc };
   enum LimiterType { none, TVB };
   enum ShockIndType { limiter, density, energy, u2 };
   static const unsigned int max_n_boundaries = 10;
   struct BoundaryConditions
   {
      int kind;
   };
   FluxType flux_type = hllc;
   LimiterType limiter_type = TVB;
   ShockIndType shock_indicator_type = limiter;
   bool char_lim = true, pos_lim = true, conserve_angular_momentum = false;
   double M = 0.0, beta = 2.0, gravity = 0.0, cfl = 0.9, time_step = -1.0;
   BoundaryConditions boundary_conditions[max_n_boundaries];
};

static int fails = 0;
#define CHECK(cond)                                                   \
   do                                                                 \
   {                                                                  \
      if (!(cond))                                                    \
      {                                                               \
         std::printf ("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
         ++fails;                                                     \
      }                                                               \
   } while (0)

int main ()
{
   const int nx = 5, ny = 3, ids[4] = {2, 1, 0, 0}; // the Sod tube's ids: left 2, right 1, walls 0
   // deal.II iterates the active cells in an order of its own: use a scrambled one
   std::vector<int> order (nx * ny);
   std::iota (order.begin (), order.end (), 0);
   for (int k = 0; k < nx * ny; ++k) std::swap (order[k], order[(7 * k + 3) % (nx * ny)]);
   dealii::Triangulation<2> tria (nx, ny, 0.0, 1.0, 0.0, 0.3, order, ids);
   dealii::DoFHandler<2> dof_handler (tria);
   // src/claw.cc:294-298: user_index = active-cell counter
   unsigned int index = 0;
   for (auto cell = dof_handler.begin_active (); cell != dof_handler.end (); ++cell) cell->set_user_index (index++);

   const unsigned int D = 4 * 6; // P2
   dflo_b200_adapter::FlatMeshStorage store;
   std::vector<uint32_t> dof_map;
   std::string why;
   CHECK (dflo_b200_adapter::flatten (dof_handler, D, store, dof_map, why) == DFLO_OK);
   const dflo_flat_mesh m = store.view ();
   CHECK (m.n_cells == nx * ny && m.n_boundary_faces == 2 * (nx + ny));

   // the library's own flattening of the same rectangle, cells numbered lexicographically
   const double args[10] = {(double) nx, (double) ny, 0.0, 1.0, 0.0, 0.3, 2, 1, 0, 0};
   dflo_mesh *ref = dflo_mesh_create ("rectangle", args, 10);
   int kinds[DFLO_MAX_BOUNDARIES], pairs[DFLO_MAX_BOUNDARIES];
   for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
   {
      kinds[b] = DFLO_BC_OUTFLOW;
      pairs[b] = -1;
   }
   CHECK (ref && dflo_mesh_flatten (ref, kinds, pairs) == 0);
   const dflo_flat_mesh *r = dflo_mesh_flat (ref);
   CHECK (r->n_cells == m.n_cells && r->n_boundary_faces == m.n_boundary_faces);
   for (int k = 0; k < m.n_cells; ++k) // adapter cell k <-> lattice cell order[k]
   {
      const int l = order[k];
      for (int d = 0; d < 2; ++d)
      {
         CHECK (std::abs (m.cell_origin[2 * k + d] - r->cell_origin[2 * l + d]) < 1e-14);
         CHECK (std::abs (m.cell_size[2 * k + d] - r->cell_size[2 * l + d]) < 1e-14);
      }
      for (int f = 0; f < 4; ++f)
      {
         const int a = m.neighbor[4 * k + f], b = r->neighbor[4 * l + f];
         CHECK ((a < 0) == (b < 0));
         if (a >= 0)
            CHECK (order[a] == b);
         else
         {
            CHECK (m.bface_cell[-1 - a] == k && m.bface_face[-1 - a] == f);
            CHECK (m.bface_id[-1 - a] == r->bface_id[-1 - b]);
         }
      }
   }
   // every interior face has exactly one owner, and it is the cell deal.II visits first
   for (int k = 0; k < m.n_cells; ++k)
      for (int f = 0; f < 4; ++f)
      {
         const int a = m.neighbor[4 * k + f];
         if (a < 0) continue;
         const bool mine = (m.face_flags[4 * k + f] & DFLO_FACE_OWNER) != 0, theirs = (m.face_flags[4 * a + (f ^ 1)] & DFLO_FACE_OWNER) != 0;
         CHECK (mine != theirs && mine == (k < a));
      }
   for (int k = 0; k < m.n_cells; ++k)
      for (unsigned int i = 0; i < D; ++i) CHECK (dof_map[(size_t) k * D + i] == (uint32_t) (k * D + i));

   // boundary expressions: read back from the ParameterHandler like parse_parameters does (src/parameters.cc:470-511)
   dealii::ParameterHandler prm;
   prm.set ("boundary_2", "w_2 value", "1.0");
   prm.set ("boundary_2", "w_3 value", "2.5*(x<0.5)");
   std::vector<std::string> expr;
   dflo_b200_adapter::capture_boundary_expressions (prm, MockAllParameters::max_n_boundaries, expr);
   CHECK (expr.size () == 40 && expr[2 * 4 + 2] == "1.0" && expr[2 * 4 + 3] == "2.5*(x<0.5)" && expr[0] == "0.0");

   MockAllParameters parameters;
   for (auto &b : parameters.boundary_conditions) b.kind = DFLO_BC_OUTFLOW;
   parameters.boundary_conditions[0].kind = DFLO_BC_SLIP;
   parameters.boundary_conditions[2].kind = DFLO_BC_INFLOW;
   dflo_params p;
   dflo_b200_adapter::fill_params (parameters, 2, false, p);
   CHECK (p.basis == DFLO_BASIS_PK && p.degree == 2 && p.flux_type == DFLO_FLUX_HLLC && p.limiter_type == DFLO_LIMITER_TVB);
   CHECK (p.bc_kind[0] == DFLO_BC_SLIP && p.bc_kind[2] == DFLO_BC_INFLOW && p.pos_lim == 1 && p.cfl == 0.9);

   // the whole path: on a GPU box a context comes back (and runs a step); without a device the library must say so
   dflo_ctx *ctx = nullptr;
   const int rc = dflo_b200_adapter::create_context (dof_handler, D, 2, false, parameters, expr, 0, &ctx, dof_map, why);
   if (rc == DFLO_OK)
   {
      std::vector<double> u ((size_t) m.n_cells * D, 0.0);
      for (int k = 0; k < m.n_cells; ++k)
      {
         u[(size_t) k * D + 2 * 6] = 1.0; // density mode 0
         u[(size_t) k * D + 3 * 6] = 2.5; // energy mode 0
      }
      CHECK (dflo_b200_set_solution (ctx, u.data (), dof_map.data (), u.size ()) == DFLO_OK);
      double t = 0.0, dt = 0.0;
      CHECK (dflo_b200_advance (ctx, 2, 1.0, &t, &dt) == DFLO_OK && t > 0.0);
      std::vector<double> v (u.size ());
      CHECK (dflo_b200_get_solution (ctx, v.data (), dof_map.data (), v.size ()) == DFLO_OK);
      double err = 0.0; // a uniform state at rest between slip walls, inflow = the same state: stays put
      for (size_t i = 0; i < u.size (); ++i) err = std::max (err, std::abs (u[i] - v[i]));
      CHECK (err < 1e-12);
      std::printf ("adapter: context created on the GPU, 2 steps to t = %g, free-stream error %.2e\n", t, err);
      dflo_b200_destroy (ctx);
   }
   else
   {
      CHECK (rc == DFLO_E_NO_DEVICE);
      std::printf ("adapter: no CUDA device here: dflo_b200_create says \"%s\"\n", why.c_str ());
   }
   dflo_mesh_destroy (ref);
   std::printf (fails ? "adapter: %d check(s) FAILED\n" : "adapter: all checks passed\n", fails);
   return fails ? 1 : 0;
}
