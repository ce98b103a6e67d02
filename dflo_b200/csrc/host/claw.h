// Host-side mirror of dflo's ConservationLaw<2> driver (reference src/claw.h:57-365) for the
// explicit path: same member names and the same control flow as src/claw.cc, with the body of
// the RK loop (assemble_system .. apply_positivity_limiter) delegated to the B200 engine through
// the C ABI of include/dflo_b200.h -- exactly the calls INTEGRATION.md asks a dflo maintainer to
// add behind `#ifdef DFLO_WITH_B200`.
#pragma once

#include "../../../include/dflo_b200.h"
#include "../tables.h"
#include "mesh.h"
#include "parameters.h"

#include <string>
#include <vector>

namespace dflo
{
   class ConservationLaw
   {
   public:
      // input_filename: input.prm; mesh_override: "<generator> <args...>" or empty (then "mesh file"
      // is read as gmsh v2, relative to the .prm); overrides: extra parameter text applied last
      ConservationLaw (const std::string &input_filename, const std::string &mesh_override, const std::string &overrides,
                       int compat);
      ~ConservationLaw ();

      bool ok () const { return error.empty (); }
      std::string error;

      // setup_system + set_initial_condition + initial limiting (src/claw.cc:981-1003)
      int setup_system (int device, int rank, int world, const void *nccl_unique_id);
      // the time loop (src/claw.cc:1026-1110), at most max_steps steps (<0: until final time)
      int run (int max_steps, bool verbose);

      int set_initial_condition (std::vector<double> &u, std::string &err) const;   // src/ic.cc:104-182
      int get_solution (std::vector<double> &u);
      int output_results (const std::string &path);                // src/output.cc:33-68 (VTU; "" or "dir/": solution-NNN.vtu + shock.vtu)
      int write_shock_file (const std::string &path);              // src/output.cc:70-79
      int compute_angular_momentum (double &value);                // src/claw.cc:604-635
      unsigned int output_file_number = 0;                         // the static counter of output.cc:47
      // run() writes the initial solution and then follows "output: time step / iter step" (src/claw.cc:1010-1017,
      // 1093-1099) into output_dir ("" = working directory like the reference) when enabled
      std::vector<std::vector<std::string>> all_files;             // the static all_files of src_mpi/output.cc:70
      int rank = 0, world = 1;                                     // set by setup_system
      bool output_enabled = false;
      std::string output_dir;
      double next_output_time = 0.0;
      int next_output_iter = 0;

      Parameters::AllParameters parameters;
      dflo_params engine_params;
      PrimitiveMesh pm;
      FlatMesh flat;
      dflo_flat_mesh flat_view;
      FeTables tab;
      dflo_ctx *ctx = nullptr;
      int compat;
      double elapsed_time = 0.0, global_dt = 0.0;
      int time_iter = 0;
      int n_dofs () const { return flat.n_cells () * tab.D; }

   private:
      void initial_value (double x, double y, double w[4]) const;
      int compute_time_step ();
      int iterate_explicit (double &res_norm0, double &res_norm, bool want_norm);
   };
}
