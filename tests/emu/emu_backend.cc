// TEST INFRASTRUCTURE ONLY -- never part of the shipped library, never loaded by dflo_b200.
//
// CPU emulation backend for Engine<Backend> (dflo_b200/csrc/engine_core.h): runs the very same
// kernel phase code as the CUDA build, block by block and thread by thread, so that index
// arithmetic, buffer rotation, limiter logic and the halo lists can be checked against the
// oracle in the CPU-only test tier before GPU time is spent.  Exposes the ABI of
// include/dflo_b200.h under the prefix dflo_emu_.  Halo exchange between emulated ranks is
// done by the test (in-process or over torch.distributed/gloo) through dflo_emu_halo_*.
#include "../../dflo_b200/csrc/abi_impl.h"

#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace
{
   struct EmuBackend
   {
      int64_t launches = 0;
      int rank = 0, world = 1;
      // halo mailboxes: outgoing[peer] = concatenated payload in send order, incoming likewise
      struct Pending { double *dst; size_t count; };
      std::map<int, std::vector<double>> outbox;
      std::map<int, std::vector<Pending>> recv_plan;
      bool halo_pending = false;

      static int64_t &cell_stage_launches ()
      {
         static int64_t n = 0;
         return n;
      }
      int open (int, int r, int w, const void *, std::string &)
      {
         rank = r;
         world = w;
         return DFLO_OK;
      }
      void close () {}
      template <class T> T *alloc (size_t n) { return static_cast<T *> (std::calloc (n ? n : 1, sizeof (T))); }
      void free (void *p) { std::free (p); }
      void h2d (void *d, const void *h, size_t b) { std::memcpy (d, h, b); }
      void d2h (void *h, const void *d, size_t b) { std::memcpy (h, d, b); }
      void zero (void *d, size_t b) { std::memset (d, 0, b); }
      void sync () {}
      int check (std::string &) { return DFLO_OK; }
      void *stream_handle () const { return nullptr; }
      void drop_graphs () {}
      bool graph_launch (int, int *) { return false; }
      bool capture_begin () { return false; }
      void capture_end_and_launch (int, int) {}
      void timer_start () {}
      void timer_stop () {}
      float timer_ms () { return 0.0f; }
      bool allreduce_min_dt (double *, bool = false, double = 0.0) { return false; } // single-rank advance only in emulation
      void allreduce_sum (double *, int) {}

      template <class K> void launch (int grid, const typename K::Args &a)
      {
         ++launches;
         std::vector<double> smem (K::SMEM_DOUBLES);
         for (int b = 0; b < grid; ++b)
            for (int p = 0; p < K::NPHASE; ++p)
               for (int t = 0; t < K::THREADS; ++t) K::phase (p, a, smem.data (), t, b);
      }
      // the stage kernel: one tile per emulated block (the CUDA build pipelines tiles through a
      // persistent kernel; the per-tile phase code is the same)
      template <class K> void launch_stage (int n_tiles, const typename K::Args &a) { launch<K> (n_tiles, a); }
      // the register-blocked Qk kernel exists only as CUDA code; the emulation runs the phase kernel
      bool use_row_kernel (int, int) const { return false; }
      int pdl_level () const { return 0; }
      bool keep_graphs_sharded () const { return false; }
      void prepare_tables (const dflo::FeTables &, const std::vector<double> &) {}
      // thread-per-cell Pk stage kernel (cell_stage.cuh): DFLO_EMU_PK=cell; default: the tile kernel
      bool use_pk_cell_kernel () const { const char *e = std::getenv ("DFLO_EMU_PK"); return e && std::string (e) == "cell"; }
      void note_cell_stage () { ++cell_stage_launches (); }
      int stage_prefetch_tiles () const { return 0; }
      int n_sms () const { return 1; }
      int debug_flags () const { return 0; }
      bool limiter_block_form () const { static const char *e = std::getenv ("DFLO_EMU_LIMITER"); return e && std::string (e) == "block"; }
      // peer-memory halo exists only on the CUDA backend
      template <class P, class V> void p2p_setup (double **, double **, const P &, const V *, int, const int *, int) {}
      bool p2p_fused_ok () const { return false; }
      const dflo::P2PFused *p2p_fused_args (int) const { return nullptr; }
      void p2p_teardown () {}
      bool p2p_exchange (int, bool = true) { return false; }
      bool p2p_deferred_ok () const { return false; }
      static bool p2p_defer_requested () { return false; }
      const dflo::P2PFused *p2p_wait_args () const { return nullptr; }
      void p2p_drain () {}
      template <class K> void launch1d (int n, const typename K::Args &a)
      {
         ++launches;
         for (int j = 0; j < n; ++j) K::thread (a, j);
      }

      // halo: sends are copied into the outbox; receives are recorded and completed by the test
      void halo_begin ()
      {
         outbox.clear ();
         recv_plan.clear ();
      }
      void halo_send (int peer, const double *buf, size_t count)
      {
         std::vector<double> &o = outbox[peer];
         o.insert (o.end (), buf, buf + count);
      }
      void halo_recv (int peer, double *dst, size_t count) { recv_plan[peer].push_back (Pending{dst, count}); }
      void halo_end () { halo_pending = !recv_plan.empty () || !outbox.empty (); }
      void halo_wait () {}
   };
}

DFLO_DEFINE_ABI (dflo_emu_, EmuBackend, dflo_emu_ctx)

extern "C" {
// number of doubles this rank sends to / expects from `peer` in the pending exchange
size_t dflo_emu_halo_send_count (dflo_emu_ctx *c, int peer) { return c->eng.bk.outbox.count (peer) ? c->eng.bk.outbox[peer].size () : 0; }
size_t dflo_emu_halo_recv_count (dflo_emu_ctx *c, int peer)
{
   size_t n = 0;
   if (c->eng.bk.recv_plan.count (peer))
      for (auto &p : c->eng.bk.recv_plan[peer]) n += p.count;
   return n;
}
void dflo_emu_halo_get_send (dflo_emu_ctx *c, int peer, double *out)
{
   const std::vector<double> &o = c->eng.bk.outbox[peer];
   std::memcpy (out, o.data (), o.size () * sizeof (double));
}
void dflo_emu_halo_put_recv (dflo_emu_ctx *c, int peer, const double *in)
{
   size_t off = 0;
   for (auto &p : c->eng.bk.recv_plan[peer])
   {
      std::memcpy (p.dst, in + off, p.count * sizeof (double));
      off += p.count;
   }
   c->eng.bk.recv_plan.erase (peer);
}
int64_t dflo_emu_cell_stage_launches (void) { return EmuBackend::cell_stage_launches (); }
// the host check that admits the thread-per-cell Pk stage kernel (cell_stage.cuh)
int dflo_emu_pk_cell_mesh_ok (const int *nbr, const unsigned char *fflags, int n_compute) { return dflo::pk_cell_mesh_ok (nbr, fflags, n_compute) ? 1 : 0; }
int dflo_emu_n_peers (dflo_emu_ctx *c) { return c->eng.lm.peers.size (); }
int dflo_emu_peer_rank (dflo_emu_ctx *c, int i) { return c->eng.lm.peers[i].rank; }
int dflo_emu_n_local (dflo_emu_ctx *c) { return c->eng.lm.n_local; }
int dflo_emu_n_compute (dflo_emu_ctx *c) { return c->eng.lm.n_compute; }
}

// Host-side check of the row-kernel tile descriptors (row_desc.h): rebuilds the local mesh the way
// the CUDA backend does for a Qk context and verifies that every face of every computed cell has
// exactly one agent, pointing at the right neighbour with the right "plus" side.  Returns the
// number of inconsistencies (0 = ok), -1 if the mesh could not be built.
extern "C" int dflo_emu_rowdesc_check (const dflo_flat_mesh *mesh, int n1, int rank, int world, int layers, int *n_tiles_out)
{
   using namespace dflo;
   LocalMesh L;
   std::string err;
   const int tc = row_tc (n1), nh = row_nh (n1);
   if (!build_local_mesh (*mesh, rank, world, layers, row_tx (n1), row_ty (n1), L, err, true)) return -1;
   if (n_tiles_out) *n_tiles_out = L.n_tiles;
   int bad = 0;
   std::vector<int> covered ((size_t) L.n_local * 4, 0);
   for (int t = 0; t < L.n_tiles; ++t)
   {
      const int *d = &L.rowdesc[(size_t) t * L.rowdesc_stride];
      const int c0 = d[0], ncb = d[1], nhl = d[2], nL = d[3], nG = d[4];
      const int *halo = d + rowd_off_halo (), *nbhi = d + rowd_off_nbhi (nh), *lj = d + rowd_off_ljob (tc, nh), *gj = d + rowd_off_gjob (tc, nh);
      if (c0 != L.tile_start[t] || ncb != L.tile_start[t + 1] - c0 || ncb > tc || nhl > nh || nL > nh || nG > nh) ++bad;
      auto su_cell = [&] (int slot) { return slot < tc ? c0 + slot : halo[slot - tc]; };
      auto expect_plus = [&] (int cell, int f) {
         const int nb = L.nbr[4 * (size_t) cell + f];
         return nb < 0 || (L.fflags[4 * (size_t) cell + f] & (DFLO_FACE_OWNER | DFLO_FACE_PERIODIC)) != 0;
      };
      for (int s = 0; s < ncb; ++s)
         for (int dir = 0; dir < 2; ++dir)
         {
            const int cell = c0 + s, f = 2 * dir + 1, code = nbhi[2 * s + dir];
            const int nb = L.nbr[4 * (size_t) cell + f];
            ++covered[4 * (size_t) cell + f];
            if (code < 0) { if (code != nb) ++bad; continue; }
            const int idx = code & 0xffff;
            const bool plus = (code & ROWD_PLUS) != 0;
            if (plus != expect_plus (cell, f)) ++bad;
            if (idx < tc)
            {
               if (c0 + idx != nb || idx >= ncb) ++bad;
               if (L.fflags[4 * (size_t) cell + f] & DFLO_FACE_PERIODIC) ++bad; // periodic partners go through ghost slots
               ++covered[4 * (size_t) nb + (f ^ 1)];
            }
            else
            {
               const int g = idx - tc;
               if (g >= nG) { ++bad; continue; }
               if (su_cell (gj[g] & 0xffff) != nb || ((gj[g] >> 16) & 1) != dir) ++bad;
               if (((gj[g] & ROWD_FLIP) != 0) != ((L.fflags[4 * (size_t) cell + f] & DFLO_FACE_FLIP) != 0)) ++bad;
            }
         }
      for (int j = 0; j < nL; ++j)
      {
         const int w = lj[2 * j], nbs = lj[2 * j + 1];
         const int s = (w & 0xffff) >> 1, dir = w & 1, cell = c0 + s, f = 2 * dir;
         const int nb = L.nbr[4 * (size_t) cell + f];
         ++covered[4 * (size_t) cell + f];
         if (s >= ncb) { ++bad; continue; }
         if (((w & ROWD_PLUS) != 0) != expect_plus (cell, f)) ++bad;
         if (nbs < 0) { if (nbs != nb) ++bad; }
         else if (su_cell (nbs) != nb) ++bad;
         if (((w & ROWD_FLIP) != 0) != ((L.fflags[4 * (size_t) cell + f] & DFLO_FACE_FLIP) != 0)) ++bad;
      }
   }
   for (int c = 0; c < L.n_compute; ++c)
      for (int f = 0; f < 4; ++f)
         if (covered[4 * (size_t) c + f] != 1) ++bad;
   return bad;
}

// point-wise device physics (dflo_b200/csrc/euler.cuh) exposed for unit tests against the oracle
extern "C" {
void dflo_emu_numerical_flux (int flux, const double n[2], const double Wp[4], const double Wm[4], const double Ap[4],
                              const double Am[4], double H[4])
{
   switch (flux)
   {
      case 0: dflo::numerical_flux<0> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
      case 1: dflo::numerical_flux<1> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
      case 2: dflo::numerical_flux<2> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
      case 3: dflo::numerical_flux<3> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
      case 4: dflo::numerical_flux<4> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
      default: dflo::numerical_flux<5> (n[0], n[1], Wp, Wm, Ap, Am, H); break;
   }
}
// flux along +e_dir of a face whose integrating ("plus") cell is the low-side one (plus_low) or the high-side one
void dflo_emu_face_flux_axis (int flux, int dir, int plus_low, const double Wl[4], const double Wr[4], const double Al[4],
                              const double Ar[4], double H[4])
{
#define DFLO_AX(F) \
   if (dir == 0) dflo::face_flux_axis<F, 0> (plus_low != 0, Wl, Wr, Al, Ar, H); \
   else dflo::face_flux_axis<F, 1> (plus_low != 0, Wl, Wr, Al, Ar, H);
   switch (flux)
   {
      case 0: DFLO_AX (0) break;
      case 1: DFLO_AX (1) break;
      case 2: DFLO_AX (2) break;
      case 3: DFLO_AX (3) break;
      case 4: DFLO_AX (4) break;
      default: DFLO_AX (5) break;
   }
#undef DFLO_AX
}
void dflo_emu_flux_x (const double W[4], double Fx[4]) { dflo::flux_x (W, Fx); }
void dflo_emu_flux_matrix (const double W[4], double F[8])
{
   double Fx[4], Fy[4];
   dflo::flux_matrix (W, Fx, Fy);
   for (int c = 0; c < 4; ++c)
   {
      F[2 * c] = Fx[c];
      F[2 * c + 1] = Fy[c];
   }
}
void dflo_emu_wminus (int kind, const double n[2], const double Wp[4], const double g[4], double Wm[4])
{
   dflo::compute_wminus (kind, n[0], n[1], Wp, g, Wm);
}
void dflo_emu_eigen (const double W[4], double Rx[16], double Lx[16], double Ry[16], double Ly[16])
{
   dflo::EigenMatrices m;
   dflo::compute_eigen_matrix (W, m);
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
      {
         Rx[4 * i + j] = m.Rx[i][j];
         Lx[4 * i + j] = m.Lx[i][j];
         Ry[4 * i + j] = m.Ry[i][j];
         Ly[4 * i + j] = m.Ly[i][j];
      }
}
double dflo_emu_minmod (double a, double b, double c, double Mdx2) { return dflo::minmod (a, b, c, Mdx2); }
void dflo_emu_eigen_stream (const double W[4], double R[16], double L[16])
{
   dflo::EigenStream m;
   dflo::compute_eigen_stream (W, m);
   for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j)
      {
         R[4 * i + j] = m.R[i][j];
         L[4 * i + j] = m.L[i][j];
      }
}
void dflo_emu_forcing_ext (const double W[4], const double f[2], double G[4]) { dflo::forcing_ext (W, f[0], f[1], G); }
}
