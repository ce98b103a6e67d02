// adapters/dealii_flatten.h -- the glue a deal.II build of cpraveen/dflo compiles in (-DDFLO_WITH_B200) to hand its
// mesh, parameters and boundary expressions to libdflo_b200 (include/dflo_b200.h).  Header-only templates over the
// deal.II types dflo already uses, so the same source compiles against deal.II 8.x/9.x and, for the syntax/ABI check
// of this repository (deal.II is not installed here), against adapters/dealii_mock -- a MOCK that provides only the
// member functions used below, with deal.II's names and signatures; it is not deal.II and proves nothing about it
// beyond "this code uses the API as documented".
//
// Call sites in the reference (INTEGRATION.md has the full diff):
//   read_parameters()   src/claw.cc:122-135   capture_boundary_expressions (prm, ...) right after parse_parameters (prm)
//   setup_system()      src/claw.cc:270-386   flatten (dof_handler, ...), fill_params (...), dflo_b200_create (...)
//   run()               src/claw.cc:989-1110  dflo_b200_set_solution / rk_stage / advance / get_solution with the dof map
#ifndef DFLO_B200_DEALII_FLATTEN_H
#define DFLO_B200_DEALII_FLATTEN_H

#include <dflo_b200.h>

#include <cstdint>
#include <string>
#include <vector>

namespace dflo_b200_adapter
{
   // storage behind a dflo_flat_mesh (the struct itself only holds pointers)
   struct FlatMeshStorage
   {
      std::vector<double> origin, size;
      std::vector<int32_t> neighbor, bface_cell, bface_face, bface_id;
      std::vector<uint8_t> face_flags, neighbor_face;
      std::vector<double> vertices;   // [nc][4][2]: what mapping = q1 works from
      std::vector<int32_t> hanging;   // [nh][6]: faces with a hanging node (coarse cell, face, fine cell 0, its face, fine cell 1, its face)
      dflo_flat_mesh view () const
      {
         dflo_flat_mesh m = dflo_flat_mesh ();
         m.n_cells = (int32_t) (origin.size () / 2);
         m.cell_origin = origin.data ();
         m.cell_size = size.data ();
         m.neighbor = neighbor.data ();
         m.face_flags = face_flags.data ();
         m.n_boundary_faces = (int32_t) bface_cell.size ();
         m.bface_cell = bface_cell.data ();
         m.bface_face = bface_face.data ();
         m.bface_id = bface_id.data ();
         m.cell_vertices = vertices.data ();
         m.neighbor_face = neighbor_face.data ();
         m.n_hanging_faces = (int32_t) (hanging.size () / 6);
         m.hanging = hanging.data ();
         return m;
      }
   };

   // Flatten once: what setup_system() computes cell by cell through deal.II iterators (neighbour arrays
   // src/claw.cc:336-380, cell numbering 294-298) becomes the SoA mesh of include/dflo_b200.h, plus the map from
   // (cell, local dof i) to deal.II's global dof index so that deal.II vectors can be handed over as they are.
   // Returns DFLO_OK, or DFLO_E_UNSUPPORTED with `why` set for what the library does not cover (with cartesian = true: cells
   // that are not axis-aligned rectangles -- pass false for mapping = q1).  Faces with hanging nodes (one level, as deal.II
   // keeps the mesh) go into FlatMeshStorage::hanging; in 2-D deal.II both cells see a line in the same direction, so
   // no DFLO_FACE_FLIP arises.
   // `cell->user_index()` must already hold the active-cell counter.
   template <class DoFHandlerType>
   int flatten (const DoFHandlerType &dof_handler, unsigned int dofs_per_cell, FlatMeshStorage &out, std::vector<uint32_t> &dof_map,
                std::string &why, bool cartesian = true)
   {
      const unsigned int nc = dof_handler.get_triangulation ().n_active_cells ();
      out.origin.assign (2 * (std::size_t) nc, 0.0);
      out.size.assign (2 * (std::size_t) nc, 0.0);
      out.neighbor.assign (4 * (std::size_t) nc, 0);
      out.face_flags.assign (4 * (std::size_t) nc, 0);
      out.neighbor_face.assign (4 * (std::size_t) nc, 0);
      out.vertices.assign (8 * (std::size_t) nc, 0.0);
      out.hanging.clear ();
      out.bface_cell.clear ();
      out.bface_face.clear ();
      out.bface_id.clear ();
      dof_map.assign ((std::size_t) nc * dofs_per_cell, 0u);
      std::vector<dealii::types::global_dof_index> idx (dofs_per_cell);
      for (typename DoFHandlerType::active_cell_iterator cell = dof_handler.begin_active (); cell != dof_handler.end (); ++cell)
      {
         const unsigned int c = cell->user_index ();
         // deal.II vertex order is lexicographic: v0 = (x0,y0), v1 = (x1,y0), v2 = (x0,y1)   
         const double x0 = cell->vertex (0)[0], y0 = cell->vertex (0)[1];
         const double hx = cell->vertex (1)[0] - x0, hy = cell->vertex (2)[1] - y0;
         const double skew = std::abs (cell->vertex (1)[1] - y0) + std::abs (cell->vertex (2)[0] - x0)
                             + std::abs (cell->vertex (3)[0] - (x0 + hx)) + std::abs (cell->vertex (3)[1] - (y0 + hy));
         for (unsigned int i = 0; i < 4; ++i)
         {
            out.vertices[8 * c + 2 * i] = cell->vertex (i)[0];
            out.vertices[8 * c + 2 * i + 1] = cell->vertex (i)[1];
         }
         if (cartesian && (!(hx > 0.0 && hy > 0.0) || skew > 1e-12 * (hx + hy)))
         {
            why = "dflo_b200: cells must be axis-aligned rectangles with local x along +x (mapping = cartesian)";
            return DFLO_E_UNSUPPORTED;
         }
         out.origin[2 * c] = x0;
         out.origin[2 * c + 1] = y0;
         out.size[2 * c] = hx;
         out.size[2 * c + 1] = hy;
         cell->get_dof_indices (idx); // FESystem of a DG element: local i = comp * n_s + node
         for (unsigned int i = 0; i < dofs_per_cell; ++i) dof_map[(std::size_t) c * dofs_per_cell + i] = (uint32_t) idx[i];
         for (unsigned int f = 0; f < 4; ++f)
         {
            if (cell->at_boundary (f))
            {
               out.neighbor[4 * c + f] = -1 - (int32_t) out.bface_cell.size ();
               out.bface_cell.push_back ((int32_t) c);
               out.bface_face.push_back ((int32_t) f);
               out.bface_id.push_back ((int32_t) cell->face (f)->boundary_id ());
               continue;
            }
            if (cell->face (f)->has_children ())
            {
               // a hanging node on this face: the two finer cells behind it, in the order of this cell's line (deal.II's
               // subface numbering).  MeshWorker integrates such a face from the fine side; the library does the same.
               out.hanging.push_back ((int32_t) c);
               out.hanging.push_back ((int32_t) f);
               for (unsigned int k = 0; k < 2; ++k)
               {
                  const typename DoFHandlerType::cell_iterator child = cell->neighbor_child_on_subface (f, k);
                  out.hanging.push_back ((int32_t) child->user_index ());
                  out.hanging.push_back ((int32_t) cell->neighbor_of_neighbor (f));
                  if (k == 0)
                  {
                     out.neighbor[4 * c + f] = (int32_t) child->user_index ();
                     out.neighbor_face[4 * c + f] = (uint8_t) cell->neighbor_of_neighbor (f);
                  }
               }
               out.face_flags[4 * c + f] = DFLO_FACE_HANGING;
               continue;
            }
            if (cell->neighbor_is_coarser (f))
            {
               // the fine side: (face, subface) of the coarser neighbour this face is a half of
               const std::pair<unsigned int, unsigned int> fs = cell->neighbor_of_coarser_neighbor (f);
               out.neighbor[4 * c + f] = (int32_t) cell->neighbor (f)->user_index ();
               out.neighbor_face[4 * c + f] = (uint8_t) fs.first;
               out.face_flags[4 * c + f] = DFLO_FACE_COARSER | DFLO_FACE_OWNER | (fs.second ? DFLO_FACE_CHILD1 : 0);
               continue;
            }
            const typename DoFHandlerType::cell_iterator nb = cell->neighbor (f);
            out.neighbor[4 * c + f] = (int32_t) nb->user_index ();
            out.neighbor_face[4 * c + f] = (uint8_t) cell->neighbor_of_neighbor (f); // f ^ 1 on lattice meshes
            // MeshWorker::loop integrates an interior face once, from the cell that compares smaller
            // (src/assemble_explicit.cc:440-451)
            if (cell < nb) out.face_flags[4 * c + f] |= DFLO_FACE_OWNER;
         }
      }
      return DFLO_OK;
   }

   // Parameters::AllParameters<2> -> dflo_params.  The enums of the reference are passed as ints: flux_type has the
   // order of src/parameters.h:229 (lxf, sw, kfvs, roe, hllc), BoundaryKind the order of src/equation.h:862-869.
   template <class AllParametersType>
   void fill_params (const AllParametersType &parameters, unsigned int degree, bool basis_is_Qk, dflo_params &p)
   {
      p = dflo_params ();
      p.basis = basis_is_Qk ? DFLO_BASIS_QK : DFLO_BASIS_PK;
      p.degree = (int32_t) degree;
      p.flux_type = (int32_t) parameters.flux_type;
      p.limiter_type = parameters.limiter_type == AllParametersType::TVB ? DFLO_LIMITER_TVB : DFLO_LIMITER_NONE;
      p.char_lim = parameters.char_lim;
      p.pos_lim = parameters.pos_lim;
      p.shock_indicator = parameters.shock_indicator_type == AllParametersType::density  ? DFLO_INDICATOR_DENSITY
                          : parameters.shock_indicator_type == AllParametersType::energy ? DFLO_INDICATOR_ENERGY
                                                                                        : DFLO_INDICATOR_LIMITER;
      p.conserve_angular_momentum = parameters.conserve_angular_momentum;
      p.M = parameters.M;
      p.beta = parameters.beta;
      p.gravity = parameters.gravity;
      p.cfl = parameters.cfl;
      p.time_step = parameters.time_step;
      p.compat = DFLO_COMPAT_SRC;
      for (unsigned int b = 0; b < AllParametersType::max_n_boundaries && b < DFLO_MAX_BOUNDARIES; ++b)
         p.bc_kind[b] = (int32_t) parameters.boundary_conditions[b].kind;
   }

   // The boundary values of input.prm are strings that parse_parameters hands straight to FunctionParser::initialize
   // (src/parameters.cc:470-511) and does not keep.  The ParameterHandler still holds them: read them again, the way
   // parse_parameters does, into expr[boundary_id * 4 + component].
   template <class ParameterHandlerType>
   void capture_boundary_expressions (ParameterHandlerType &prm, unsigned int max_n_boundaries, std::vector<std::string> &expr)
   {
      expr.assign ((std::size_t) max_n_boundaries * 4, "0.0");
      for (unsigned int b = 0; b < max_n_boundaries; ++b)
      {
         prm.enter_subsection ("boundary_" + dealii::Utilities::int_to_string (b));
         for (unsigned int c = 0; c < 4; ++c) expr[(std::size_t) b * 4 + c] = prm.get ("w_" + dealii::Utilities::int_to_string (c) + " value");
         prm.leave_subsection ();
      }
   }

   // flatten + create + boundary expressions in one call; returns the library's error code (text: dflo_b200_last_error)
   template <class DoFHandlerType, class AllParametersType>
   int create_context (const DoFHandlerType &dof_handler, unsigned int dofs_per_cell, unsigned int degree, bool basis_is_Qk,
                       const AllParametersType &parameters, const std::vector<std::string> &bc_expr, int device, dflo_ctx **ctx,
                       std::vector<uint32_t> &dof_map, std::string &why)
   {
      FlatMeshStorage store;
      int rc = flatten (dof_handler, dofs_per_cell, store, dof_map, why);
      if (rc != DFLO_OK) return rc;
      dflo_params p;
      fill_params (parameters, degree, basis_is_Qk, p);
      const dflo_flat_mesh m = store.view ();
      rc = dflo_b200_create (&m, &p, device, ctx);
      if (rc != DFLO_OK)
      {
         why = dflo_b200_last_error (nullptr);
         return rc;
      }
      for (unsigned int b = 0; b < DFLO_MAX_BOUNDARIES && (std::size_t) b * 4 + 3 < bc_expr.size (); ++b)
         for (unsigned int c = 0; c < 4; ++c)
         {
            rc = dflo_b200_set_boundary_expression (*ctx, (int) b, (int) c, bc_expr[(std::size_t) b * 4 + c].c_str ());
            if (rc != DFLO_OK)
            {
               why = dflo_b200_last_error (*ctx);
               return rc;
            }
         }
      return DFLO_OK;
   }
}
#endif
