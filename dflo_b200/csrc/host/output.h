// VTU writers of the output path (reference src/output.cc:33-87, Postprocessor src/equation.cc:59-187).
// Pure host code: the solution arrives as the host copy of current_solution in the reference DoF layout.
#pragma once

#include "../tables.h"
#include "mesh.h"

#include <string>

namespace dflo
{
   // solution-NNN.vtu: what DataOut::build_patches (mapping, fe.degree) + write_vtu produce for a DG field --
   // every cell is cut into max(degree,1)^2 sub-quads whose vertices carry the cell's own polynomial, so the
   // field stays discontinuous across cells.  Point data: XMomentum YMomentum Density Energy XVelocity
   // YVelocity Pressure [schlieren_plot].  Returns false when the file cannot be written.
   bool write_solution_vtu (const FeTables &tab, const FlatMesh &flat, const double *u, bool schlieren_plot, double time,
                            unsigned int cycle, const std::string &path);
   // shock.vtu: one quad per cell, cell data mu_shock (null: zeros) and shock_indicator (src/output.cc:70-79)
   bool write_shock_vtu (const FlatMesh &flat, const double *mu_shock, const double *shock_indicator, const std::string &path);
}
