// VTU writers of the output path (reference src/output.cc:33-87, Postprocessor src/equation.cc:59-187).
// Pure host code: the solution arrives as the host copy of current_solution in the reference DoF layout.
#pragma once

#include "../tables.h"
#include "mesh.h"

#include <string>
#include <vector>

namespace dflo
{
   // solution-NNN.vtu: what DataOut::build_patches (mapping, fe.degree) + write_vtu produce for a DG field --
   // every cell is cut into max(degree,1)^2 sub-quads whose vertices carry the cell's own polynomial, so the
   // field stays discontinuous across cells.  Point data as DataOutBase::write_vtu orders them: the vector ranges
   // XMomentum__YMomentum and XVelocity__YVelocity (3 components, z = 0), then the scalars Density Energy Pressure
   // [schlieren_plot] [subdomain].  Returns false when the file cannot be written.
   // [cell_begin, cell_end) (cell_end < 0: all cells): the piece one process of the MPI tree writes, with
   // subdomain >= 0 adding the "subdomain" array of src_mpi/output.cc:51-54; u is always the global vector.
   bool write_solution_vtu (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, double time,
                            unsigned int cycle, const std::string &path, int cell_begin = 0, int cell_end = -1, int subdomain = -1);
   // solution-NNN.plt for "output: format = tecplot" (src/output.cc:51-52, 65-66): the same patches and variables as
   // an ASCII FEBLOCK zone of quadrilaterals
   bool write_solution_tecplot (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, double time,
                                const std::string &path);
   bool write_shock_tecplot (const FlatMesh &flat, const double *mu_shock, const double *shock_indicator, const std::string &path);
   // compute_angular_momentum (src/claw.cc:604-635): sum over cells [cell_begin, cell_end) of int (x m_y - y m_x)
   double angular_momentum (const FeTables &tab, const FlatMesh &flat, const double *u, int cell_begin = 0, int cell_end = -1);
   // master_file.visit (DataOutBase::write_visit_record, src_mpi/output.cc:70-84): "!NBLOCKS n" and then the n piece
   // files of every output so far, one per line
   bool write_visit_record (const std::vector<std::vector<std::string>> &all_files, const std::string &path);
   // shock.vtu: one quad per cell, mu_shock (null: zeros) and shock_indicator (src/output.cc:70-79) -- written as point
   // data constant on each cell's four vertices, which is how DataOut writes cell vectors
   bool write_shock_vtu (const FlatMesh &flat, const double *mu_shock, const double *shock_indicator, const std::string &path);
}
