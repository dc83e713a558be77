// TEST INFRASTRUCTURE ONLY -- definitions behind the t8mini declarations in t8.h: a single-process, single-rank
// facade with the t8code / libsc / MPI call signatures the t8gpu reference uses, implemented on oracle/miniforest.c
// (Cartesian one-tree quad/hex forests in Morton order).  It lets the reference's OWN MeshManager /
// SubgridMeshManager / solver code run unmodified so that the oracle and the product can be pinned against it.
// Semantics follow SURVEY.md App. C; this is not t8code and is "parity unpinned" against real t8code.
#include <t8.h>

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

extern "C" {
// oracle/miniforest.c
struct mf_forest;
mf_forest* mf_new_uniform(int dim, int level, int periodic);
void       mf_free(mf_forest*);
int64_t    mf_num_elements(const mf_forest*);
int        mf_dim(const mf_forest*);
int        mf_face_neighbors(const mf_forest*, int64_t e, int face, int64_t out[4]);
void       mf_element(const mf_forest*, int64_t e, int* level, uint32_t coord[3]);
int        mf_is_family(const mf_forest*, int64_t i, const int64_t* rank_offsets, int nranks);
mf_forest* mf_apply_adapt(const mf_forest*, const int8_t* res);
}

static const int MAXL = 20;

struct t8mini_element {
  int      level;
  uint32_t c[3];
  int      dim;
};
struct t8mini_cmesh {
  int dim, periodic;
};
struct t8mini_scheme {
  t8_eclass_scheme_c quad, hex;
};
struct t8mini_forest {
  mf_forest*                  f = nullptr;
  std::vector<t8mini_element> elems;
  int                         refcount = 1, committed = 0, dim = 0;
  void*                       user_data = nullptr;
  t8mini_scheme*              scheme    = nullptr;
  // pending construction
  t8_forest_t       from = nullptr;
  t8_forest_adapt_t adapt_fn = nullptr;
  int               do_partition = 0;
};

static t8mini_scheme g_scheme = {{2}, {3}};

static void fill_elems(t8_forest_t F) {
  int64_t n = mf_num_elements(F->f);
  F->elems.resize(n);
  for (int64_t i = 0; i < n; i++) {
    mf_element(F->f, i, &F->elems[i].level, F->elems[i].c);
    F->elems[i].dim = F->dim;
  }
}

// ---- scheme
int t8_eclass_scheme_c::t8_element_level(const t8_element_t* e) { return e->level; }
int t8_eclass_scheme_c::t8_element_num_faces(const t8_element_t* e) { return 2 * e->dim; }
int t8_eclass_scheme_c::t8_element_child_id(const t8_element_t* e) {
  if (e->level == 0) return 0;
  int id = 0;
  for (int d = 0; d < e->dim; d++) id |= (int)((e->c[d] >> (MAXL - e->level)) & 1u) << d;
  return id;
}
void t8_eclass_scheme_c::t8_element_destroy(int n, t8_element_t** e) {
  for (int i = 0; i < n; i++) delete e[i];
}

extern "C" {

// ---- MPI / sc (one rank)
int MPI_Comm_size(MPI_Comm, int* s) { *s = 1; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int* r) { *r = 0; return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
int MPI_Allgather(const void* sb, int sc, MPI_Datatype, void* rb, int rc, MPI_Datatype, MPI_Comm) {
  if (sb != MPI_IN_PLACE) memcpy(rb, sb, (size_t)(sc > 0 ? sc : rc));
  return MPI_SUCCESS;
}
int MPI_Allreduce(const void* sb, void* rb, int count, MPI_Datatype dt, MPI_Op, MPI_Comm) {
  size_t es = dt == MPI_DOUBLE ? 8 : (dt == MPI_FLOAT || dt == MPI_INT ? 4 : 1);
  if (sb != MPI_IN_PLACE) memcpy(rb, sb, es * count);
  return MPI_SUCCESS;
}
sc_array* sc_array_new_data(void* base, size_t es, size_t ec) {
  sc_array* a = new sc_array{base, es, ec};
  return a;
}
void sc_array_destroy(sc_array* a) { delete a; }
int  sc_MPI_Init(int*, char***) { return MPI_SUCCESS; }
int  sc_MPI_Finalize(void) { return MPI_SUCCESS; }
void sc_init(sc_MPI_Comm, int, int, void*, int) {}
void sc_finalize(void) {}
void t8_init(int) {}

// ---- cmesh / scheme
t8_scheme_cxx_t* t8_scheme_new_default_cxx(void) { return &g_scheme; }
t8_cmesh_t t8_cmesh_new_periodic(sc_MPI_Comm, int dim) { return new t8mini_cmesh{dim, 1}; }
t8_cmesh_t t8mini_cmesh_new_cube(int dim, int periodic) { return new t8mini_cmesh{dim, periodic}; }
t8_cmesh_t t8_cmesh_new_prismed_spherical_shell_icosahedron(double, double, int, int, sc_MPI_Comm) {
  SC_ABORT("t8mini: prism/spherical-shell cmesh needs real t8code");
  return nullptr;
}
void t8_cmesh_destroy(t8_cmesh_t* c) {
  delete *c;
  *c = nullptr;
}

// ---- forest life cycle
t8_forest_t t8_forest_new_uniform(t8_cmesh_t cmesh, t8_scheme_cxx_t* scheme, int level, int, sc_MPI_Comm) {
  t8_forest_t F = new t8mini_forest();
  F->dim        = cmesh->dim;
  F->f          = mf_new_uniform(cmesh->dim, level, cmesh->periodic);
  F->scheme     = scheme;
  F->committed  = 1;
  fill_elems(F);
  return F;
}
void t8_forest_init(t8_forest_t* pF) { *pF = new t8mini_forest(); }
void t8_forest_ref(t8_forest_t F) { F->refcount++; }
void t8_forest_unref(t8_forest_t* pF) {
  t8_forest_t F = *pF;
  if (--F->refcount == 0) {
    mf_free(F->f);
    delete F;
  }
  *pF = nullptr;
}
int  t8_forest_is_committed(t8_forest_t F) { return F->committed; }
void t8_forest_set_adapt(t8_forest_t F, t8_forest_t from, t8_forest_adapt_t fn, int recursive) {
  if (recursive) SC_ABORT("t8mini: recursive adapt not supported");
  F->from     = from;
  F->adapt_fn = fn;
}
void t8_forest_set_ghost(t8_forest_t, int, t8_ghost_type_t) {}
void t8_forest_set_balance(t8_forest_t F, t8_forest_t from, int) {
  if (from && !F->from) F->from = from;
}
void t8_forest_set_partition(t8_forest_t F, t8_forest_t from, int) {
  F->from         = from;
  F->do_partition = 1;
}
void t8_forest_set_user_data(t8_forest_t F, void* d) { F->user_data = d; }
void* t8_forest_get_user_data(t8_forest_t F) { return F->user_data; }

void t8_forest_commit(t8_forest_t F) {
  t8_forest_t from = F->from;
  if (!from) SC_ABORT("t8mini: commit without a source forest");
  F->dim    = from->dim;
  F->scheme = from->scheme;
  int64_t n = mf_num_elements(from->f);
  std::vector<int8_t> res(n, 0);
  if (F->adapt_fn) {
    // t8code's non-recursive adapt: walk the leaves; offer a family when the next 2^d leaves form one
    t8_eclass_scheme_c* ts  = F->dim == 3 ? &from->scheme->hex : &from->scheme->quad;
    int                 nch = 1 << F->dim;
    std::vector<t8_element_t*> ptrs(nch);
    for (int64_t i = 0; i < n;) {
      int fam = mf_is_family(from->f, i, nullptr, 1);
      int num = fam ? nch : 1;
      for (int c = 0; c < num; c++) ptrs[c] = &from->elems[i + c];
      int r = F->adapt_fn(F, from, 0, (t8_locidx_t)i, ts, fam, num, ptrs.data());
      if (r < 0 && !fam) r = 0;
      res[i] = (int8_t)(r > 0 ? 1 : (r < 0 ? -1 : 0));
      i += r < 0 ? nch : 1;
    }
  }
  F->f = mf_apply_adapt(from->f, res.data());  // identity + (no-op) balance when nothing is flagged
  fill_elems(F);
  F->committed = 1;
  // set_adapt / set_partition consume one reference of the source forest (the reference t8_forest_ref()s before)
  t8_forest_t tmp = from;
  t8_forest_unref(&tmp);
  F->from = nullptr;
}

// ---- queries
t8_locidx_t t8_forest_get_local_num_elements(t8_forest_t F) { return (t8_locidx_t)mf_num_elements(F->f); }
t8_locidx_t t8_forest_get_num_ghosts(t8_forest_t) { return 0; }
t8_locidx_t t8_forest_get_num_local_trees(t8_forest_t) { return 1; }
t8_eclass_t t8_forest_get_tree_class(t8_forest_t F, t8_locidx_t) { return F->dim == 3 ? T8_ECLASS_HEX : T8_ECLASS_QUAD; }
t8_eclass_scheme_c* t8_forest_get_eclass_scheme(t8_forest_t F, t8_eclass_t c) {
  return c == T8_ECLASS_HEX ? &F->scheme->hex : &F->scheme->quad;
}
t8_locidx_t t8_forest_get_tree_num_elements(t8_forest_t F, t8_locidx_t) { return (t8_locidx_t)mf_num_elements(F->f); }
t8_locidx_t t8_forest_get_tree_element_offset(t8_forest_t, t8_locidx_t) { return 0; }
t8_element_t* t8_forest_get_element_in_tree(t8_forest_t F, t8_locidx_t, t8_locidx_t i) { return &F->elems[i]; }

double t8_forest_element_volume(t8_forest_t, t8_locidx_t, const t8_element_t* e) {
  double h = std::ldexp(1.0, -e->level);
  return e->dim == 3 ? h * h * h : h * h;
}
double t8_forest_element_face_area(t8_forest_t, t8_locidx_t, const t8_element_t* e, int) {
  double h = std::ldexp(1.0, -e->level);
  return e->dim == 3 ? h * h : h;
}
void t8_forest_element_face_normal(t8_forest_t, t8_locidx_t, const t8_element_t*, int face, double n[3]) {
  n[0] = n[1] = n[2] = 0.0;
  n[face >> 1]       = (face & 1) ? 1.0 : -1.0;
}
void t8_forest_element_centroid(t8_forest_t, t8_locidx_t, const t8_element_t* e, double* x) {
  double h = std::ldexp(1.0, -e->level);
  for (int d = 0; d < 3; d++) x[d] = d < e->dim ? std::ldexp((double)e->c[d], -MAXL) + 0.5 * h : 0.0;
}

void t8_forest_leaf_face_neighbors(t8_forest_t F, t8_locidx_t, const t8_element_t* leaf, t8_element_t** pn[], int face,
                                   int* dual_faces[], int* num_neighbors, t8_locidx_t** pidx,
                                   t8_eclass_scheme_c** pscheme, int) {
  int64_t e = leaf - F->elems.data();
  int64_t nb[4];
  int     nn = mf_face_neighbors(F->f, e, face, nb);
  *num_neighbors = nn;
  *pscheme       = F->dim == 3 ? &F->scheme->hex : &F->scheme->quad;
  *pn            = (t8_element_t**)malloc(sizeof(t8_element_t*) * (nn ? nn : 1));
  *dual_faces    = (int*)malloc(sizeof(int) * (nn ? nn : 1));
  *pidx          = (t8_locidx_t*)malloc(sizeof(t8_locidx_t) * (nn ? nn : 1));
  for (int i = 0; i < nn; i++) {
    (*pn)[i]         = new t8mini_element(F->elems[nb[i]]);
    (*dual_faces)[i] = face ^ 1;
    (*pidx)[i]       = (t8_locidx_t)nb[i];
  }
}

void t8_forest_ghost_exchange_data(t8_forest_t, sc_array*) {}
void t8_forest_partition_data(t8_forest_t, t8_forest_t, const sc_array* in, sc_array* out) {
  memcpy(out->array, in->array, in->elem_size * in->elem_count);
}
// No file is written (t8code's VTK writer is not restated); the call is captured so that tests can check WHAT the
// managers hand to t8code: the forest (element count, levels) and the data fields.
struct t8mini_vtk_capture {
  int                 calls = 0, num_data = 0, dim = 0, min_level = 0, max_level = 0;
  int64_t             n_elements = 0;
  std::string         prefix;
  std::vector<int>    types;
  std::vector<std::string>         names;
  std::vector<std::vector<double>> data;
};
static t8mini_vtk_capture g_vtk;

int t8_forest_write_vtk_ext(t8_forest_t F, const char* prefix, int, int, int, int, int, int, int, int num_data,
                            t8_vtk_data_field_t* fields) {
  g_vtk.calls++;
  g_vtk.prefix     = prefix ? prefix : "";
  g_vtk.num_data   = num_data;
  g_vtk.dim        = F->dim;
  g_vtk.n_elements = mf_num_elements(F->f);
  g_vtk.min_level = MAXL; g_vtk.max_level = 0;
  for (auto const& e : F->elems) {
    g_vtk.min_level = e.level < g_vtk.min_level ? e.level : g_vtk.min_level;
    g_vtk.max_level = e.level > g_vtk.max_level ? e.level : g_vtk.max_level;
  }
  g_vtk.types.clear(); g_vtk.names.clear(); g_vtk.data.clear();
  for (int k = 0; k < num_data; k++) {
    const int comps = fields[k].type == T8_VTK_VECTOR ? 3 : 1;
    g_vtk.types.push_back((int)fields[k].type);
    g_vtk.names.emplace_back(fields[k].description);
    g_vtk.data.emplace_back(fields[k].data, fields[k].data + (size_t)comps * g_vtk.n_elements);
  }
  return 1;
}
// test access to the capture: info = calls, num_data, n_elements, dim, min_level, max_level
void t8mini_vtk_info(int64_t info[6]) {
  info[0] = g_vtk.calls; info[1] = g_vtk.num_data; info[2] = g_vtk.n_elements; info[3] = g_vtk.dim;
  info[4] = g_vtk.min_level; info[5] = g_vtk.max_level;
}
const char* t8mini_vtk_prefix(void) { return g_vtk.prefix.c_str(); }
const char* t8mini_vtk_field_name(int k) { return k < (int)g_vtk.names.size() ? g_vtk.names[k].c_str() : ""; }
int64_t t8mini_vtk_field(int k, double* out, int64_t cap) {   // returns the number of doubles of field k (copies <= cap)
  if (k < 0 || k >= (int)g_vtk.data.size()) return -1;
  const int64_t n = (int64_t)g_vtk.data[k].size();
  if (out) memcpy(out, g_vtk.data[k].data(), sizeof(double) * (size_t)(n < cap ? n : cap));
  return n;
}
}  // extern "C"
