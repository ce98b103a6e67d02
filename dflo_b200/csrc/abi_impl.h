// Bodies of the C ABI of include/dflo_b200.h, generic over the backend.  engine_cuda.cu
// instantiates them as dflo_b200_* for the product; tests/emu instantiates the same bodies as
// dflo_emu_* over the CPU emulation backend (test infrastructure).
#pragma once

#include "engine_core.h"

#define DFLO_ABI_CAT2(a, b) a##b
#define DFLO_ABI_CAT(a, b) DFLO_ABI_CAT2 (a, b)

// PREFIX: symbol prefix (dflo_b200_ or dflo_emu_); BACKEND: backend type; CTX: opaque struct name
#define DFLO_DEFINE_ABI(PREFIX, BACKEND, CTX)                                                                            \
   struct CTX                                                                                                            \
   {                                                                                                                     \
      dflo::Engine<BACKEND> eng;                                                                                         \
   };                                                                                                                    \
   static std::string DFLO_ABI_CAT (PREFIX, create_error_text);                                                          \
   static void DFLO_ABI_CAT (PREFIX, set_create_error) (const char *s) { DFLO_ABI_CAT (PREFIX, create_error_text) = s; } \
   static const char *DFLO_ABI_CAT (PREFIX, get_create_error) () { return DFLO_ABI_CAT (PREFIX, create_error_text).c_str (); } \
   extern "C" {                                                                                                          \
   int DFLO_ABI_CAT (PREFIX, abi_version) (void) { return DFLO_B200_ABI_VERSION; }                                       \
   const char *DFLO_ABI_CAT (PREFIX, strerror) (int code)                                                                \
   {                                                                                                                     \
      switch (code)                                                                                                      \
      {                                                                                                                  \
         case DFLO_OK: return "ok";                                                                                      \
         case DFLO_E_INVALID: return "invalid argument";                                                                 \
         case DFLO_E_UNSUPPORTED: return "unsupported configuration";                                                    \
         case DFLO_E_CUDA: return "CUDA error";                                                                          \
         case DFLO_E_NEGATIVE_STATE: return "Fatal: Negative states";                                                    \
         case DFLO_E_POSLIM_ROOT: return "Problem in positivity limiter";                                                \
         case DFLO_E_NCCL: return "NCCL error";                                                                          \
         case DFLO_E_EXPR: return "boundary expression syntax error";                                                    \
         case DFLO_E_NO_DEVICE: return "no CUDA device available (the engine has no CPU fallback)";                      \
         default: return "unknown error";                                                                                \
      }                                                                                                                  \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, create_sharded) (const dflo_flat_mesh *mesh, const dflo_params *prm, int device, int rank,  \
                                              int world, const void *nccl_id, CTX **out)                                 \
   {                                                                                                                     \
      if (!mesh || !prm || !out) return DFLO_E_INVALID;                                                                  \
      *out = nullptr;                                                                                                    \
      CTX *c = new CTX;                                                                                                  \
      int rc = c->eng.bk.open (device, rank, world, nccl_id, c->eng.error);                                              \
      if (rc == DFLO_OK) rc = c->eng.init (*mesh, *prm, rank, world);                                                    \
      if (rc != DFLO_OK)                                                                                                 \
      {                                                                                                                  \
         DFLO_ABI_CAT (PREFIX, set_create_error) (c->eng.error.c_str ());                                                \
         c->eng.bk.close ();                                                                                             \
         delete c;                                                                                                       \
         return rc;                                                                                                      \
      }                                                                                                                  \
      *out = c;                                                                                                          \
      return DFLO_OK;                                                                                                    \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, create) (const dflo_flat_mesh *mesh, const dflo_params *prm, int device, CTX **out)         \
   {                                                                                                                     \
      return DFLO_ABI_CAT (PREFIX, create_sharded) (mesh, prm, device, 0, 1, nullptr, out);                              \
   }                                                                                                                     \
   void DFLO_ABI_CAT (PREFIX, destroy) (CTX *c)                                                                          \
   {                                                                                                                     \
      if (!c) return;                                                                                                    \
      c->eng.release ();                                                                                                 \
      c->eng.bk.close ();                                                                                                \
      delete c;                                                                                                          \
   }                                                                                                                     \
   const char *DFLO_ABI_CAT (PREFIX, last_error) (const CTX *c)                                                          \
   {                                                                                                                     \
      return c ? c->eng.error.c_str () : DFLO_ABI_CAT (PREFIX, get_create_error) ();                                     \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, dofs_per_cell) (const CTX *c) { return c->eng.tab.D; }                                      \
   int DFLO_ABI_CAT (PREFIX, n_q_face) (const CTX *c) { return c->eng.tab.n1; }                                          \
   int DFLO_ABI_CAT (PREFIX, n_rk) (const CTX *c) { return c->eng.n_rk; }                                                \
   double DFLO_ABI_CAT (PREFIX, ark) (const CTX *c, int rk) { return (rk >= 0 && rk < c->eng.n_rk) ? c->eng.ark[rk] : 0.0; } \
   int64_t DFLO_ABI_CAT (PREFIX, n_cells_owned) (const CTX *c) { return c->eng.lm.n_owned; }                             \
   int64_t DFLO_ABI_CAT (PREFIX, cell_range) (const CTX *c, int64_t *b, int64_t *e)                                      \
   {                                                                                                                     \
      if (b) *b = c->eng.lm.begin;                                                                                       \
      if (e) *e = c->eng.lm.end;                                                                                         \
      return c->eng.lm.n_owned;                                                                                          \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, set_solution) (CTX *c, const double *u, const uint32_t *m, size_t n)                        \
   {                                                                                                                     \
      return (c && u) ? c->eng.set_solution (u, m, n) : DFLO_E_INVALID;                                                  \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, get_solution) (CTX *c, double *u, const uint32_t *m, size_t n)                              \
   {                                                                                                                     \
      return (c && u) ? c->eng.get_solution (u, m, n) : DFLO_E_INVALID;                                                  \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, get_cell_average) (CTX *c, double *a) { return (c && a) ? c->eng.get_cell_average (a) : DFLO_E_INVALID; } \
   int DFLO_ABI_CAT (PREFIX, commit_step) (CTX *c) { return c ? c->eng.commit_step () : DFLO_E_INVALID; }                \
   int DFLO_ABI_CAT (PREFIX, set_boundary_values) (CTX *c, const double *g)                                              \
   {                                                                                                                     \
      return (c && g) ? c->eng.set_boundary_values (g) : DFLO_E_INVALID;                                                 \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, set_boundary_expression) (CTX *c, int id, int comp, const char *e)                          \
   {                                                                                                                     \
      return (c && e) ? c->eng.set_boundary_expression (id, comp, e) : DFLO_E_INVALID;                                   \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, set_external_force) (CTX *c, const char *fx, const char *fy)                                \
   {                                                                                                                     \
      return (c && fx && fy) ? c->eng.set_external_force (fx, fy) : DFLO_E_INVALID;                                      \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, assemble_rhs) (CTX *c, double t) { return c ? c->eng.assemble_rhs (t) : DFLO_E_INVALID; }   \
   int DFLO_ABI_CAT (PREFIX, get_rhs) (CTX *c, double *r, const uint32_t *m, size_t n)                                   \
   {                                                                                                                     \
      return (c && r) ? c->eng.get_rhs (r, m, n) : DFLO_E_INVALID;                                                       \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, rk_stage) (CTX *c, int rk, double t, double dt, double *res)                                \
   {                                                                                                                     \
      return c ? c->eng.rk_stage (rk, t, dt, res) : DFLO_E_INVALID;                                                      \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, compute_dt) (CTX *c, double t, double tf, double *dt)                                       \
   {                                                                                                                     \
      return (c && dt) ? c->eng.compute_dt (t, tf, dt) : DFLO_E_INVALID;                                                 \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, limit_initial_condition) (CTX *c) { return c ? c->eng.limit_initial_condition () : DFLO_E_INVALID; } \
   int DFLO_ABI_CAT (PREFIX, advance) (CTX *c, int n, double tf, double *t, double *dt)                                  \
   {                                                                                                                     \
      return (c && t) ? c->eng.advance (n, tf, t, dt) : DFLO_E_INVALID;                                                  \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, poll_error) (CTX *c) { return c ? c->eng.poll_error () : DFLO_E_INVALID; }                  \
   int DFLO_ABI_CAT (PREFIX, get_limited_flags) (CTX *c, int32_t *f) { return (c && f) ? c->eng.get_limited_flags (f) : DFLO_E_INVALID; } \
   int DFLO_ABI_CAT (PREFIX, get_shock_indicator) (CTX *c, double *s) { return (c && s) ? c->eng.get_shock_indicator (s) : DFLO_E_INVALID; } \
   int64_t DFLO_ABI_CAT (PREFIX, launch_count) (const CTX *c) { return c ? c->eng.bk.launches : 0; }                     \
   void *DFLO_ABI_CAT (PREFIX, stream) (const CTX *c) { return c ? c->eng.bk.stream_handle () : nullptr; }               \
   int DFLO_ABI_CAT (PREFIX, synchronize) (CTX *c)                                                                       \
   {                                                                                                                     \
      if (!c) return DFLO_E_INVALID;                                                                                     \
      c->eng.bk.sync ();                                                                                                 \
      return c->eng.bk.check (c->eng.error);                                                                             \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, time_stage_kernel) (CTX *c, int rk, int reps, size_t flush_bytes, float *ms)                \
   {                                                                                                                     \
      return (c && ms) ? c->eng.time_stage_kernel (rk, reps, flush_bytes, ms) : DFLO_E_INVALID;                          \
   }                                                                                                                     \
   int DFLO_ABI_CAT (PREFIX, last_advance_ms) (CTX *c, float *ms)                                                        \
   {                                                                                                                     \
      if (!c || !ms) return DFLO_E_INVALID;                                                                              \
      *ms = c->eng.bk.timer_ms ();                                                                                       \
      return DFLO_OK;                                                                                                    \
   }                                                                                                                     \
   }
