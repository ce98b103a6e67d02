// Host-side mesh handling of the standalone driver: what dflo gets from gmsh + deal.II's
// GridIn::read_msh / Triangulation (reference src/claw.cc:956-967) and flattens once in
// setup_system (src/claw.cc:270-386).
#pragma once

#include "../../../include/dflo_b200.h"

#include <string>
#include <vector>

namespace dflo
{
   // gmsh-like primitive mesh: vertices, quads in deal.II lexicographic vertex order
   // (v0=(0,0) v1=(1,0) v2=(0,1) v3=(1,1)), boundary lines with physical ids.
   struct PrimitiveMesh
   {
      std::vector<double> vertices;  // [nv][2]
      std::vector<int> cells;        // [nc][4]
      std::vector<int> blines;       // [nb][2]
      std::vector<int> bline_id;     // [nb]
      int n_vertices () const { return vertices.size () / 2; }
      int n_cells () const { return cells.size () / 4; }
      int n_blines () const { return bline_id.size (); }
   };

   // Owning storage behind a dflo_flat_mesh view.
   struct FlatMesh
   {
      std::vector<double> origin, size;
      std::vector<int32_t> neighbor;
      std::vector<uint8_t> face_flags;
      std::vector<int32_t> bface_cell, bface_face, bface_id;
      std::vector<double> vertices;        // [nc][4][2]
      std::vector<uint8_t> neighbor_face;  // [nc][4]
      std::vector<int32_t> hanging;        // [nh][6] faces with a hanging node (dflo_flat_mesh::hanging)
      bool cartesian = true;               // every cell an axis-aligned rectangle, every neighbour across face f on its face f ^ 1
      std::string why_not_cartesian;
      dflo_flat_mesh view () const;
      // the cell's map from the unit square (bilinear through the four vertices; on Cartesian cells origin + xi * size),
      // and its Jacobian J = d(x,y)/d(xi,eta) as {x_xi, x_eta, y_xi, y_eta}
      void map (int cell, double xi, double eta, double &x, double &y) const
      {
         const double *q = &vertices[8 * (size_t) cell];
         const double n0 = (1.0 - xi) * (1.0 - eta), n1 = xi * (1.0 - eta), n2 = (1.0 - xi) * eta, n3 = xi * eta;
         x = n0 * q[0] + n1 * q[2] + n2 * q[4] + n3 * q[6];
         y = n0 * q[1] + n1 * q[3] + n2 * q[5] + n3 * q[7];
      }
      void jacobian (int cell, double xi, double eta, double J[4]) const
      {
         const double *q = &vertices[8 * (size_t) cell];
         J[0] = (q[2] - q[0]) * (1.0 - eta) + (q[6] - q[4]) * eta;
         J[1] = (q[4] - q[0]) * (1.0 - xi) + (q[6] - q[2]) * xi;
         J[2] = (q[3] - q[1]) * (1.0 - eta) + (q[7] - q[5]) * eta;
         J[3] = (q[5] - q[1]) * (1.0 - xi) + (q[7] - q[3]) * xi;
      }
      int n_cells () const { return origin.size () / 2; }
      int n_bfaces () const { return bface_cell.size (); }
   };

   // Derive neighbours, MeshWorker face ownership, periodic partners and the boundary-face list.
   // bc_kind[id] == DFLO_BC_PERIODIC marks periodic ids, periodic_pair[id] their partner id.
   // General straight-sided quadrilaterals are accepted (mapping = q1): FlatMesh::cartesian says whether the mesh also
   // qualifies for mapping = cartesian.  Returns false (and sets err) for cells with a non-positive Jacobian,
   // non-manifold faces or unmatched periodic faces.
   bool flatten (const PrimitiveMesh &pm, const int bc_kind[DFLO_MAX_BOUNDARIES],
                 const int periodic_pair[DFLO_MAX_BOUNDARIES], FlatMesh &out, std::string &err);

   // structured nx x ny block of [x0,x1]x[y0,y1], cells x fastest; ids of (left,right,bottom,top)
   PrimitiveMesh make_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4]);

   // the same rectangle with smoothly displaced interior vertices (general quadrilaterals for mapping = q1); rotate: mixed cell orientations
   PrimitiveMesh make_skewed_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4], double amp, int rotate);

   // a rectangle of nx x ny cells in which the cells i0 <= i < i1, j0 <= j < j1 are split into four: hanging nodes on the rim of the patch
   PrimitiveMesh make_refined_rectangle (int nx, int ny, double x0, double x1, double y0, double y1, const int ids[4], int i0, int i1, int j0, int j1,
                                         int rotate = 0);
   PrimitiveMesh make_compression_corner (int nx1_cells, int nx2_cells, int ny_cells);   // examples/compression_corner/corner.geo (mapping = q1)

   // The four BASELINE geometries, reproducing the transfinite blocks of the reference's .geo
   // files (no gmsh in this image).  Cells are emitted block by block like gmsh does.
   PrimitiveMesh make_isentropic_vortex_grid (int n_cells_per_side);          // examples/isentropic_vortex/grid.geo
   PrimitiveMesh make_sod_tube (int nx_cells, int ny_cells);                  // examples/sod_shock_tube/tube.geo
   PrimitiveMesh make_double_mach_grid (int ny_cells);                        // examples/double_mach_reflection/grid.geo
   PrimitiveMesh make_forward_step_grid (double cl);                          // examples/forward_step/step.geo

   // gmsh ASCII format 2 (what GridIn::read_msh reads)
   bool read_gmsh2 (const std::string &path, PrimitiveMesh &pm, std::string &err);
   bool write_gmsh2 (const std::string &path, const PrimitiveMesh &pm);
}
