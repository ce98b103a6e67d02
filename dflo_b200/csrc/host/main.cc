// dflo_b200: standalone driver with dflo's command line (reference src/main.cc:13-81):
//     dflo_b200 input.prm [--mesh "<generator> <args>"] [--steps N] [--compat src|mpi] [--vtu out.vtu] [--output-dir DIR] [--quiet]
// Reads the same input.prm, prints the same per-step lines (src/claw.cc:1031-1041, 768) and runs
// the explicit RK stages on GPU 0 through the C ABI.  --output-dir DIR turns on the reference's output schedule
// (initial solution, "output: time step / iter step", final time; solution-NNN.vtu + shock.vtu, src/output.cc)
// with the files going to DIR ("." = the working directory, which is where the reference writes).
#include "../../../include/dflo_host.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

int main (int argc, char **argv)
{
   if (argc < 2)
   {
      std::fprintf (stderr, "usage: %s input.prm [--mesh \"kind args\"] [--steps N] [--compat src|mpi] [--vtu file] [--output-dir DIR] [--quiet]\n", argv[0]);
      return 1;
   }
   std::string mesh, vtu, outdir;
   bool output = false;
   int steps = -1, compat = DFLO_COMPAT_SRC, verbose = 1;
   for (int i = 2; i < argc; ++i)
   {
      if (!std::strcmp (argv[i], "--mesh") && i + 1 < argc) mesh = argv[++i];
      else if (!std::strcmp (argv[i], "--steps") && i + 1 < argc) steps = std::atoi (argv[++i]);
      else if (!std::strcmp (argv[i], "--compat") && i + 1 < argc) compat = std::strcmp (argv[++i], "mpi") ? DFLO_COMPAT_SRC : DFLO_COMPAT_MPI;
      else if (!std::strcmp (argv[i], "--vtu") && i + 1 < argc) vtu = argv[++i];
      else if (!std::strcmp (argv[i], "--output-dir") && i + 1 < argc)
      {
         output = true;
         outdir = argv[++i];
      }
      else if (!std::strcmp (argv[i], "--quiet")) verbose = 0;
   }
   const auto t0 = std::chrono::steady_clock::now ();
   dflo_claw *claw = dflo_claw_create (argv[1], mesh.empty () ? nullptr : mesh.c_str (), nullptr, compat);
   if (!claw)
   {
      std::fprintf (stderr, "\n----------------------------------------------------\nException on processing: %s\nAborting!\n"
                            "----------------------------------------------------\n", dflo_host_last_error ());
      return 1;
   }
   if (output) dflo_claw_set_output (claw, outdir == "." ? "" : outdir.c_str ());
   int rc = dflo_claw_setup (claw, 0, 0, 1, nullptr);
   double t = 0.0;
   int done = 0;
   if (!rc) rc = dflo_claw_run (claw, steps, verbose, &t, &done);
   if (!rc && !vtu.empty ()) rc = dflo_claw_write_vtu (claw, vtu.c_str ());
   if (rc)
   {
      std::fprintf (stderr, "\n----------------------------------------------------\nException on processing: %s (%s)\nAborting!\n"
                            "----------------------------------------------------\n", dflo_host_last_error (), dflo_b200_strerror (rc));
      dflo_claw_destroy (claw);
      return 1;
   }
   const double wall = std::chrono::duration<double> (std::chrono::steady_clock::now () - t0).count ();
   std::printf ("\n%d steps, T = %g, %d dofs; Elapsed wall time = %g min\n", done, t, dflo_claw_n_dofs (claw), wall / 60.0);
   dflo_claw_destroy (claw);
   return 0;
}
