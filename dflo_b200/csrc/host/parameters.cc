#include "parameters.h"

#include <cstdlib>
#include <fstream>
#include <sstream>

namespace dflo
{
   namespace Parameters
   {
      namespace
      {
         std::string trim (const std::string &s)
         {
            size_t b = 0, e = s.size ();
            while (b < e && std::isspace ((unsigned char) s[b])) ++b;
            while (e > b && std::isspace ((unsigned char) s[e - 1])) --e;
            return s.substr (b, e - b);
         }
         // collapse runs of blanks inside names ("diffusion power       = 2.0")
         std::string squeeze (const std::string &s)
         {
            std::string o;
            bool sp = false;
            for (char ch : trim (s))
            {
               if (std::isspace ((unsigned char) ch))
                  sp = true;
               else
               {
                  if (sp && !o.empty ()) o += ' ';
                  sp = false;
                  o += ch;
               }
            }
            return o;
         }
         bool to_bool (const std::string &v, bool &out)
         {
            if (v == "true" || v == "yes" || v == "on" || v == "1") { out = true; return true; }
            if (v == "false" || v == "no" || v == "off" || v == "0") { out = false; return true; }
            return false;
         }
      }

      // Parameters::*::declare_parameters, src/parameters.cc:10-419
      AllParameters::AllParameters ()
      {
         auto d = [&] (const std::string &k, const std::string &v) { entries[k] = v; };
         d ("mesh type", "gmsh");
         d ("mesh file", "grid.msh");
         d ("degree", "1");
         d ("basis", "Qk");
         d ("mapping", "q1");
         d ("diffusion power", "2.0");
         d ("diffusion coefficient", "0.0");
         d ("gravity", "0.0");
         d ("f_0 value", "0.0"); // components of the external force, src_mpi/parameters.cc:355-360
         d ("f_1 value", "0.0");
         d ("time stepping/stationary", "false");
         d ("time stepping/cfl", "0.0");
         d ("time stepping/time step type", "global");
         d ("time stepping/time step", "-1.0");
         d ("time stepping/final time", "1.0e20");
         d ("time stepping/theta scheme value", "1.0");
         d ("time stepping/nonlinear iterations", "1");
         for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
         {
            const std::string s = "boundary_" + std::to_string (b) + "/";
            d (s + "type", "outflow");
            for (int c = 0; c < 4; ++c) d (s + "w_" + std::to_string (c) + " value", "0.0");
            d (s + "pair", "0");        // src_mpi/parameters.cc:407
            d (s + "direction", "x");   // src_mpi/parameters.cc:409
         }
         d ("initial condition/function", "none");
         for (int c = 0; c < 4; ++c) d ("initial condition/w_" + std::to_string (c) + " value", "0.0");
         d ("linear solver/output", "quiet");
         d ("linear solver/method", "rk3");
         d ("linear solver/residual", "1e-10");
         d ("linear solver/max iters", "300");
         d ("linear solver/ilut fill", "2");
         d ("linear solver/ilut absolute tolerance", "1e-9");
         d ("linear solver/ilut relative tolerance", "1.1");
         d ("linear solver/ilut drop tolerance", "1e-10");
         d ("refinement/refinement", "true");
         d ("refinement/time step", "1.0e20");
         d ("refinement/iter step", "100000000");
         d ("refinement/refinement fraction", "0.1");
         d ("refinement/unrefinement fraction", "0.1");
         d ("refinement/max elements", "1000000");
         d ("refinement/shock value", "4.0");
         d ("refinement/shock levels", "3.0");
         d ("flux/flux", "lxf");
         d ("flux/stab", "mesh");
         d ("flux/stab value", "1");
         d ("limiter/shock indicator", "limiter");
         d ("limiter/type", "none");
         d ("limiter/characteristic limiter", "false");
         d ("limiter/positivity limiter", "false");
         d ("limiter/M", "0");
         d ("limiter/beta", "1.0");
         d ("limiter/conserve angular momentum", "false");
         d ("output/schlieren plot", "false");
         d ("output/time step", "1e20");
         d ("output/iter step", "1000000");
         d ("output/format", "vtk");
         d ("output/compute angular momentum", "10000000");
      }

      // ParameterHandler text format
      bool AllParameters::parse_text (const std::string &text, std::string &err)
      {
         std::istringstream in (text);
         std::string line;
         std::vector<std::string> path;
         int lineno = 0;
         while (std::getline (in, line))
         {
            ++lineno;
            const size_t hash = line.find ('#');
            if (hash != std::string::npos) line = line.substr (0, hash);
            line = trim (line);
            if (line.empty ()) continue;
            if (line.compare (0, 10, "subsection") == 0 && (line.size () == 10 || std::isspace ((unsigned char) line[10])))
            {
               path.push_back (squeeze (line.substr (10)));
               continue;
            }
            if (line == "end")
            {
               if (path.empty ())
               {
                  err = "line " + std::to_string (lineno) + ": 'end' without subsection";
                  return false;
               }
               path.pop_back ();
               continue;
            }
            if (line.compare (0, 3, "set") == 0 && line.size () > 3 && std::isspace ((unsigned char) line[3]))
            {
               const size_t eq = line.find ('=');
               if (eq == std::string::npos)
               {
                  err = "line " + std::to_string (lineno) + ": missing '='";
                  return false;
               }
               std::string key;
               for (auto &p : path) key += p + "/";
               key += squeeze (line.substr (4, eq - 4));
               auto it = entries.find (key);
               if (it == entries.end ())
               {
                  err = "line " + std::to_string (lineno) + ": no such entry was declared: " + key;
                  return false;
               }
               it->second = trim (line.substr (eq + 1));
               continue;
            }
            err = "line " + std::to_string (lineno) + ": cannot parse '" + line + "'";
            return false;
         }
         if (!path.empty ())
         {
            err = "unclosed subsection " + path.back ();
            return false;
         }
         return true;
      }

      bool AllParameters::parse_file (const std::string &path, std::string &err)
      {
         std::ifstream in (path.c_str ());
         if (!in)
         {
            err = "cannot open " + path;
            return false;
         }
         std::stringstream ss;
         ss << in.rdbuf ();
         return parse_text (ss.str (), err);
      }

      // AllParameters::parse_parameters, src/parameters.cc:422-551
      bool AllParameters::finish (std::string &err)
      {
         auto get = [&] (const std::string &k) { return entries[k]; };
         auto num = [&] (const std::string &k) { return std::atof (entries[k].c_str ()); };
         auto sel = [&] (const std::string &k, const std::string &choices) {
            const std::string v = "|" + entries[k] + "|";
            if (("|" + choices + "|").find (v) == std::string::npos)
            {
               err = "entry '" + k + "' = '" + entries[k] + "' does not match " + choices;
               return false;
            }
            return true;
         };
         if (!sel ("mesh type", "ucd|gmsh") || !sel ("basis", "Qk|Pk") || !sel ("mapping", "q1|q2|cartesian")
             || !sel ("time stepping/time step type", "global|local") || !sel ("linear solver/output", "quiet|verbose")
             || !sel ("linear solver/method", "gmres|direct|umfpack|rk3|mood") || !sel ("flux/flux", "lxf|sw|kfvs|roe|hllc|kep")
             || !sel ("flux/stab", "constant|mesh") || !sel ("limiter/shock indicator", "limiter|density|energy|u2")
             || !sel ("limiter/type", "none|TVB|minmax") || !sel ("output/format", "vtk|tecplot")
             || !sel ("initial condition/function", "none|rt|isenvort|vortsys"))
            return false;
         mesh_type = get ("mesh type");
         mesh_filename = get ("mesh file");
         degree = (int) num ("degree");
         basis = get ("basis");
         mapping = get ("mapping");
         diffusion_power = num ("diffusion power");
         diffusion_coef = num ("diffusion coefficient");
         gravity = num ("gravity");
         external_force[0] = get ("f_0 value"); // src_mpi/parameters.cc:488-497
         external_force[1] = get ("f_1 value");
         cfl = num ("time stepping/cfl");
         time_step_type = get ("time stepping/time step type");
         time_step = num ("time stepping/time step");
         final_time = num ("time stepping/final time");
         if (!to_bool (get ("time stepping/stationary"), is_stationary))
         {
            err = "stationary must be a boolean";
            return false;
         }
         if (is_stationary)
         {
            time_step = 1.0;
            final_time = 1.0e20;
         }
         else if (!(cfl > 0 || time_step > 0))
         {
            err = "cfl and time_step zero";
            return false;
         }
         theta = num ("time stepping/theta scheme value");
         max_nonlin_iter = (int) num ("time stepping/nonlinear iterations");
         for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
         {
            const std::string s = "boundary_" + std::to_string (b) + "/";
            if (!sel (s + "type", "slip|inflow|outflow|pressure|farfield|periodic")) return false;
            boundary_type[b] = get (s + "type");
            for (int c = 0; c < 4; ++c) boundary_expr[b][c] = get (s + "w_" + std::to_string (c) + " value");
            periodic_pair[b] = boundary_type[b] == "periodic" ? (int) num (s + "pair") : -1;
            periodic_direction[b] = get (s + "direction");
         }
         ic_function = get ("initial condition/function");
         for (int c = 0; c < 4; ++c) ic_expr[c] = get ("initial condition/w_" + std::to_string (c) + " value");
         solver_output = get ("linear solver/output");
         solver_method = get ("linear solver/method");
         bool ok = to_bool (get ("refinement/refinement"), do_refine) && to_bool (get ("limiter/characteristic limiter"), char_lim)
                   && to_bool (get ("limiter/positivity limiter"), pos_lim)
                   && to_bool (get ("limiter/conserve angular momentum"), conserve_angular_momentum)
                   && to_bool (get ("output/schlieren plot"), schlieren_plot);
         if (!ok)
         {
            err = "boolean entry expected true|false";
            return false;
         }
         flux = get ("flux/flux");
         shock_indicator = get ("limiter/shock indicator");
         limiter_type = get ("limiter/type");
         M = num ("limiter/M");
         beta = num ("limiter/beta");
         if (beta < 1.0 || beta > 2.0)
         {
            err = "limiter beta must lie in [1,2]";
            return false;
         }
         output_time_step = num ("output/time step");
         output_iter_step = (int) num ("output/iter step");
         output_format = get ("output/format");
         ang_mom_step = (int) num ("output/compute angular momentum");

         // cross-checks, src/parameters.cc:536-550
         if (solver_method == "mood" && time_step_type != "global")
         {
            err = "MOOD requires global time step";
            return false;
         }
         if (solver_method == "mood" && basis != "Pk")
         {
            err = "MOOD is implemented only for Pk";
            return false;
         }
         if (limiter_type == "TVB" && mapping != "cartesian")
         {
            err = "TVB limiter works on cartesian grids only";
            return false;
         }
         if (limiter_type == "minmax" && basis != "Qk") // src_mpi/parameters.cc:610-611
         {
            err = "minmax limiter is implemented only for Qk";
            return false;
         }
         if (basis == "Pk" && mapping != "cartesian")
         {
            err = "Pk basis can only be used with Cartesian grids";
            return false;
         }
         if (basis == "Pk" && do_refine)
         {
            err = "Refinement does not work for Pk basis";
            return false;
         }
         return true;
      }

      bool AllParameters::to_engine_params (int compat, dflo_params &p, std::string &err) const
      {
         // what the B200 engine covers: explicit SSP-RK on Cartesian cells
         if (solver_method != "rk3")
         {
            err = "only 'method = rk3' (explicit SSP-RK) runs on the B200 engine; got " + solver_method;
            return false;
         }
         if (mapping != "cartesian" && mapping != "q1")
         {
            err = "only 'mapping = cartesian' and 'mapping = q1' run on the B200 engine; got " + mapping;
            return false;
         }
         if (time_step_type == "local" && !(cfl > 0.0))
         {
            err = "time step type = local needs a cfl";
            return false;
         }
         if (do_refine)
         {
            err = "mesh refinement is not supported by the B200 engine (set refinement = false)";
            return false;
         }
         if (shock_indicator == "u2")
         {
            err = "'shock indicator = u2' belongs to the MOOD path, which the B200 engine does not cover";
            return false;
         }
         if (diffusion_coef != 0.0)
         {
            err = "artificial viscosity is not part of the explicit path (diffusion coefficient must be 0)";
            return false;
         }
         p.basis = basis == "Qk" ? DFLO_BASIS_QK : DFLO_BASIS_PK;
         p.degree = degree;
         p.flux_type = flux == "lxf" ? DFLO_FLUX_LXF : flux == "sw" ? DFLO_FLUX_SW : flux == "kfvs" ? DFLO_FLUX_KFVS
                     : flux == "roe" ? DFLO_FLUX_ROE : flux == "kep" ? DFLO_FLUX_KEP : DFLO_FLUX_HLLC;
         p.limiter_type = limiter_type == "TVB" ? DFLO_LIMITER_TVB : limiter_type == "minmax" ? DFLO_LIMITER_MINMAX : DFLO_LIMITER_NONE;
         p.char_lim = char_lim;
         p.pos_lim = pos_lim;
         p.conserve_angular_momentum = conserve_angular_momentum;
         p.compat = compat;
         p.mapping = mapping == "q1" ? DFLO_MAPPING_Q1 : DFLO_MAPPING_CARTESIAN; // claw.cc:165-190
         p.local_time_step = time_step_type == "local"; // claw.cc:456, 469, 709
         p.reserved1 = 0;
         p.shock_indicator = shock_indicator == "density" ? DFLO_INDICATOR_DENSITY : shock_indicator == "energy" ? DFLO_INDICATOR_ENERGY : DFLO_INDICATOR_LIMITER; // parameters.cc:229-237
         p.M = M;
         p.beta = beta;
         p.gravity = gravity;
         p.cfl = cfl;
         p.time_step = time_step;
         for (int b = 0; b < DFLO_MAX_BOUNDARIES; ++b)
         {
            const std::string &t = boundary_type[b];
            p.bc_kind[b] = t == "slip" ? DFLO_BC_SLIP : t == "inflow" ? DFLO_BC_INFLOW : t == "pressure" ? DFLO_BC_PRESSURE
                         : t == "farfield" ? DFLO_BC_FARFIELD : t == "periodic" ? DFLO_BC_PERIODIC : DFLO_BC_OUTFLOW;
         }
         return true;
      }
   }
}
