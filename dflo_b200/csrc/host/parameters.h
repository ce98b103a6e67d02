// input.prm of dflo: deal.II ParameterHandler syntax with the schema the reference declares in
// src/parameters.cc:10-419 (plus the periodic boundary keys of src_mpi/parameters.cc:397-410).
// Same names, same defaults, same cross-checks (parameters.cc:536-550); undeclared keys are
// errors, like in deal.II.
#pragma once

#include "../../../include/dflo_b200.h"

#include <map>
#include <string>
#include <vector>

namespace dflo
{
   namespace Parameters
   {
      struct AllParameters
      {
         // top level
         std::string mesh_type, mesh_filename;
         int degree;
         std::string basis;          // "Qk" | "Pk"
         std::string mapping;        // "q1" | "q2" | "cartesian"
         double diffusion_power, diffusion_coef, gravity;
         std::string external_force[2]; // "f_0 value", "f_1 value" in x,y,t (MPI tree only)
         // time stepping
         bool is_stationary;
         double cfl, time_step, final_time, theta;
         std::string time_step_type; // "global" | "local"
         int max_nonlin_iter;
         // boundaries
         std::string boundary_type[DFLO_MAX_BOUNDARIES];
         std::string boundary_expr[DFLO_MAX_BOUNDARIES][4];
         int periodic_pair[DFLO_MAX_BOUNDARIES];
         std::string periodic_direction[DFLO_MAX_BOUNDARIES];
         // initial condition
         std::string ic_function;    // none | rt | isenvort | vortsys
         std::string ic_expr[4];
         // linear solver
         std::string solver_output, solver_method;
         // refinement
         bool do_refine;
         // flux / limiter
         std::string flux;
         std::string shock_indicator, limiter_type;
         bool char_lim, pos_lim, conserve_angular_momentum;
         double M, beta;
         // output
         bool schlieren_plot;
         double output_time_step;
         int output_iter_step;
         std::string output_format;
         int ang_mom_step;

         // every declared entry with its current value, key = "subsection/name" ("" for top level)
         std::map<std::string, std::string> entries;

         AllParameters ();                       // declare_parameters: defaults
         bool parse_text (const std::string &text, std::string &err);   // ParameterHandler::read_input
         bool parse_file (const std::string &path, std::string &err);
         bool finish (std::string &err);         // parse_parameters: typed fields + cross-checks
         // the subset the device engine needs
         bool to_engine_params (int compat, dflo_params &out, std::string &err) const;
      };
   }
}
