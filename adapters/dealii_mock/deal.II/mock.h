// MOCK of the handful of deal.II classes adapters/dealii_flatten.h touches -- NOT deal.II.  It exists so that the
// adapter is compiled and exercised in this repository's CPU test tier (deal.II is not installed in the image): same
// class / member names and signatures as deal.II 8.x-9.x for exactly the calls the adapter makes, on a structured
// nx x ny mesh whose active cells are numbered in an arbitrary order and whose DoFs are numbered cell by cell.
#ifndef DFLO_B200_DEALII_MOCK_H
#define DFLO_B200_DEALII_MOCK_H

#include <cmath>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace dealii
{
   namespace types
   {
      typedef unsigned int global_dof_index;
      typedef unsigned char boundary_id;
   }
   namespace Utilities
   {
      inline std::string int_to_string (unsigned int i) { return std::to_string (i); }
   }
   template <int dim>
   struct Point
   {
      double c[dim];
      double operator[] (unsigned int i) const { return c[i]; }
   };

   // ParameterHandler: the subset parse_parameters uses (enter_subsection / get / leave_subsection)
   class ParameterHandler
   {
   public:
      void set (const std::string &section, const std::string &key, const std::string &value) { values[section + "/" + key] = value; }
      void enter_subsection (const std::string &s) { path = s; }
      void leave_subsection () { path.clear (); }
      std::string get (const std::string &key) const
      {
         const auto it = values.find (path + "/" + key);
         return it == values.end () ? std::string ("0.0") : it->second;
      }

   private:
      std::map<std::string, std::string> values;
      std::string path;
   };

   template <int dim> class DoFHandler;

   template <int dim>
   class Triangulation
   {
   public:
      // structured mesh of [x0,x1] x [y0,y1]; `order[k]` = lattice index (i + nx j) of the k-th active cell, i.e. the
      // iteration order of begin_active() (deal.II does not promise a lexicographic one); boundary ids: left, right, bottom, top
      Triangulation (int nx_, int ny_, double x0_, double x1_, double y0_, double y1_, const std::vector<int> &order_, const int ids_[4])
         : nx (nx_), ny (ny_), x0 (x0_), y0 (y0_), hx ((x1_ - x0_) / nx_), hy ((y1_ - y0_) / ny_), order (order_), rank (order_.size ())
      {
         for (int k = 0; k < 4; ++k) ids[k] = ids_[k];
         for (std::size_t k = 0; k < order.size (); ++k) rank[order[k]] = (int) k;
         user.assign (order.size (), 0u);
      }
      unsigned int n_active_cells () const { return (unsigned int) order.size (); }
      int nx, ny;
      double x0, y0, hx, hy;
      std::vector<int> order, rank;
      int ids[4];
      mutable std::vector<unsigned int> user;
   };

   template <int dim>
   class DoFHandler
   {
   public:
      struct FaceAccessor
      {
         types::boundary_id id;
         bool children;
         types::boundary_id boundary_id () const { return id; }
         bool has_children () const { return children; }
         const FaceAccessor *operator-> () const { return this; }
      };
      // plays TriaIterator and its accessor at once: `cell->f()` and `cell < other` both work
      struct cell_iterator
      {
         const DoFHandler *dh;
         int k; // position in the active-cell order
         const cell_iterator *operator-> () const { return this; }
         cell_iterator &operator++ ()
         {
            ++k;
            return *this;
         }
         bool operator!= (const cell_iterator &o) const { return k != o.k; }
         bool operator< (const cell_iterator &o) const { return k < o.k; } // same level: ordered by index
         int lattice () const { return dh->tria->order[k]; }
         unsigned int user_index () const { return dh->tria->user[k]; }
         void set_user_index (unsigned int v) const { dh->tria->user[k] = v; }
         Point<dim> vertex (unsigned int v) const
         {
            const Triangulation<dim> &t = *dh->tria;
            const int i = lattice () % t.nx, j = lattice () / t.nx;
            Point<dim> p;
            p.c[0] = t.x0 + t.hx * (i + (v & 1));
            p.c[1] = t.y0 + t.hy * (j + (v >> 1));
            return p;
         }
         bool at_boundary (unsigned int f) const
         {
            const Triangulation<dim> &t = *dh->tria;
            const int i = lattice () % t.nx, j = lattice () / t.nx;
            return (f == 0 && i == 0) || (f == 1 && i == t.nx - 1) || (f == 2 && j == 0) || (f == 3 && j == t.ny - 1);
         }
         FaceAccessor face (unsigned int f) const
         {
            FaceAccessor a;
            a.id = (types::boundary_id) dh->tria->ids[f];
            a.children = false;
            return a;
         }
         bool neighbor_is_coarser (unsigned int) const { return false; }
         unsigned int neighbor_of_neighbor (unsigned int f) const { return f ^ 1u; }
         // refinement: the mock mesh has none -- these exist so that the adapter's hanging-node branch compiles
         cell_iterator neighbor_child_on_subface (unsigned int f, unsigned int) const { return neighbor (f); }
         std::pair<unsigned int, unsigned int> neighbor_of_coarser_neighbor (unsigned int f) const { return std::make_pair (f ^ 1u, 0u); }
         cell_iterator neighbor (unsigned int f) const
         {
            const Triangulation<dim> &t = *dh->tria;
            const int l = lattice () + (f == 0 ? -1 : f == 1 ? 1 : f == 2 ? -t.nx : t.nx);
            cell_iterator n = {dh, t.rank[l]};
            return n;
         }
         void get_dof_indices (std::vector<types::global_dof_index> &idx) const
         {
            for (std::size_t i = 0; i < idx.size (); ++i) idx[i] = (types::global_dof_index) (k * idx.size () + i);
         }
      };
      typedef cell_iterator active_cell_iterator;
      explicit DoFHandler (const Triangulation<dim> &t) : tria (&t) {}
      const Triangulation<dim> &get_triangulation () const { return *tria; }
      active_cell_iterator begin_active () const
      {
         active_cell_iterator c = {this, 0};
         return c;
      }
      active_cell_iterator end () const
      {
         active_cell_iterator c = {this, (int) tria->n_active_cells ()};
         return c;
      }
      const Triangulation<dim> *tria;
   };
}
#endif
