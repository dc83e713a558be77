/* TEST INFRASTRUCTURE ONLY -- "mini-forest": a CPU restatement of the t8code behaviour that the t8gpu hot path
 * consumes, for Cartesian (one-tree, unit square / unit cube, optionally periodic) quad / hex forests, plus a
 * restatement of the reference's host-side connectivity construction on top of it.
 *
 * t8code itself (un-vendored, un-pinned dependency of the reference: CMakeLists.txt:33 `find_package(T8CODE REQUIRED)`,
 * API shape ~ v2.0) is NOT available in this image, so the forest semantics restated here follow t8code's published
 * behaviour as the reference's call sites assume it:
 *   - leaves in Morton order, x = bit 0, y = bit 1, z = bit 2 of the child id (subgrid_mesh_manager.inl:616-631)
 *   - face ids f0=-x f1=+x f2=-y f3=+y f4=-z f5=+z (subgrid_mesh_manager.inl:590-599)
 *   - t8_forest_leaf_face_neighbors: 1 neighbour (same size or coarser), 2^(d-1) (finer, SFC order), 0 at a boundary
 *   - partition: contiguous SFC ranges, first element of rank p = floor(N*p/P)
 *   - ghosts: face neighbours owned by other ranks, ordered by owner rank then SFC index; local id = N_local + g
 *   - adapt: non-recursive, callback contract of mesh_manager.inl:125-162, followed by 2:1 face balance
 * "parity unpinned" against real t8code; pinned against the reference's own compute_connectivity_information code
 * through oracle/_ref (t8mini), see oracle/README.md.
 *
 * Reference code restated:
 *   t8gpu/mesh/mesh_manager.inl:332-481            MeshManager::compute_connectivity_information
 *   t8gpu/mesh/subgrid_mesh_manager.inl:560-961    add_face + SubgridMeshManager::compute_connectivity_information
 *   t8gpu/mesh/mesh_manager.inl:125-162,258-281    adapt callback, old->new element map
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXL 20

typedef struct {
  int       dim;
  int       periodic;
  int64_t   n;
  uint64_t* key;   /* Morton key of the anchor at MAXL resolution */
  int8_t*   level;
} mf_forest;

static uint64_t spread(uint32_t v, int dim) {
  uint64_t r = 0;
  for (int b = 0; b < MAXL; b++) r |= (uint64_t)((v >> b) & 1u) << (dim * b);
  return r;
}
static uint32_t compact(uint64_t k, int dim) {
  uint32_t r = 0;
  for (int b = 0; b < MAXL; b++) r |= (uint32_t)((k >> (dim * b)) & 1u) << b;
  return r;
}
static uint64_t morton(const uint32_t c[3], int dim) {
  uint64_t k = 0;
  for (int d = 0; d < dim; d++) k |= spread(c[d], dim) << d;
  return k;
}
static void unmorton(uint64_t k, int dim, uint32_t c[3]) {
  c[0] = c[1] = c[2] = 0;
  for (int d = 0; d < dim; d++) c[d] = compact(k >> d, dim);
}

mf_forest* mf_new_uniform(int dim, int level, int periodic) {
  mf_forest* f = (mf_forest*)calloc(1, sizeof(mf_forest));
  f->dim       = dim;
  f->periodic  = periodic;
  f->n         = (int64_t)1 << (dim * level);
  f->key       = (uint64_t*)malloc(sizeof(uint64_t) * f->n);
  f->level     = (int8_t*)malloc(f->n);
  for (int64_t i = 0; i < f->n; i++) {
    f->key[i]   = (uint64_t)i << (dim * (MAXL - level));
    f->level[i] = (int8_t)level;
  }
  return f;
}
void mf_free(mf_forest* f) {
  if (!f) return;
  free(f->key);
  free(f->level);
  free(f);
}
int64_t mf_num_elements(const mf_forest* f) { return f->n; }
int     mf_dim(const mf_forest* f) { return f->dim; }

/* levels, centroid (3 doubles, z = 0 in 2-D), volume (t8_forest_element_volume: h^dim), child id */
void mf_get_elements(const mf_forest* f, int32_t* levels, double* centroids, double* volumes, int32_t* child_ids) {
  for (int64_t i = 0; i < f->n; i++) {
    uint32_t c[3];
    unmorton(f->key[i], f->dim, c);
    int    l = f->level[i];
    double h = ldexp(1.0, -l);
    if (levels) levels[i] = l;
    if (centroids)
      for (int d = 0; d < 3; d++) centroids[3 * i + d] = d < f->dim ? ldexp((double)c[d], -MAXL) + 0.5 * h : 0.0;
    if (volumes) volumes[i] = f->dim == 3 ? h * h * h : h * h;
    if (child_ids) child_ids[i] = l == 0 ? 0 : (int32_t)((f->key[i] >> (f->dim * (MAXL - l))) & ((1u << f->dim) - 1));
  }
}

/* index of the leaf containing Morton key k */
static int64_t find_leaf(const mf_forest* f, uint64_t k) {
  int64_t lo = 0, hi = f->n - 1;
  while (lo < hi) {
    int64_t mid = (lo + hi + 1) >> 1;
    if (f->key[mid] <= k)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

/* t8_forest_leaf_face_neighbors restated (global leaf indices). Returns the number of neighbours (0,1,2^(d-1)). */
static int face_neighbors(const mf_forest* f, int64_t e, int face, int64_t out[4]) {
  int      dim = f->dim, l = f->level[e], ax = face >> 1, sgn = (face & 1) ? 1 : -1;
  uint32_t c[3];
  unmorton(f->key[e], dim, c);
  uint32_t h    = 1u << (MAXL - l);
  uint32_t full = 1u << MAXL;
  int64_t  nc   = (int64_t)c[ax] + (sgn > 0 ? (int64_t)h : -(int64_t)h);
  if (nc < 0 || nc >= (int64_t)full) {
    if (!f->periodic) return 0;
    nc = (nc + full) % full;
  }
  uint32_t a[3] = {c[0], c[1], c[2]};
  a[ax]         = (uint32_t)nc;
  int64_t idx   = find_leaf(f, morton(a, dim));
  if (f->level[idx] <= l) {
    out[0] = idx;
    return 1;
  }
  /* finer: children of the same-size neighbour that touch the shared face, in SFC (child id) order */
  uint32_t hh = h >> 1;
  int      nn = 0;
  for (int ch = 0; ch < (1 << dim); ch++) {
    int bit = (ch >> ax) & 1;
    /* neighbour on +side: its children with bit 0 along ax touch us; on -side: bit 1 */
    if (bit != (sgn > 0 ? 0 : 1)) continue;
    uint32_t b[3] = {a[0], a[1], a[2]};
    for (int d = 0; d < dim; d++)
      if ((ch >> d) & 1) b[d] += hh;
    out[nn++] = find_leaf(f, morton(b, dim));
  }
  return nn;
}

/* 2:1 face balance: refine any leaf that has a face neighbour more than one level finer, to a fixed point.
 * flags_out semantics are internal. */
static mf_forest* refine_marked(const mf_forest* f, const uint8_t* mark) {
  int64_t nn = 0;
  for (int64_t i = 0; i < f->n; i++) nn += mark[i] ? (1 << f->dim) : 1;
  mf_forest* g = (mf_forest*)calloc(1, sizeof(mf_forest));
  g->dim       = f->dim;
  g->periodic  = f->periodic;
  g->n         = nn;
  g->key       = (uint64_t*)malloc(sizeof(uint64_t) * nn);
  g->level     = (int8_t*)malloc(nn);
  int64_t o    = 0;
  for (int64_t i = 0; i < f->n; i++) {
    if (mark[i]) {
      int l = f->level[i] + 1;
      for (int ch = 0; ch < (1 << f->dim); ch++) {
        g->key[o]   = f->key[i] | ((uint64_t)ch << (f->dim * (MAXL - l)));
        g->level[o] = (int8_t)l;
        o++;
      }
    } else {
      g->key[o]   = f->key[i];
      g->level[o] = f->level[i];
      o++;
    }
  }
  return g;
}

static mf_forest* balance(mf_forest* f) {
  /* Seen from the fine side: a leaf of level l whose same-size face neighbour position lies inside a leaf of
   * level < l-1 forces that leaf to be refined.  Iterate to the (unique, minimal) fixed point. */
  for (;;) {
    uint8_t* mark = (uint8_t*)calloc(f->n, 1);
    int      any  = 0;
    for (int64_t e = 0; e < f->n; e++) {
      int l = f->level[e];
      if (l < 2) continue;
      for (int face = 0; face < 2 * f->dim; face++) {
        int64_t nb[4];
        int     nn = face_neighbors(f, e, face, nb);
        if (nn == 1 && f->level[nb[0]] < l - 1) {
          mark[nb[0]] = 1;
          any         = 1;
        }
      }
    }
    if (!any) {
      free(mark);
      return f;
    }
    mf_forest* g = refine_marked(f, mark);
    free(mark);
    mf_free(f);
    f = g;
  }
}

/* exported views used by oracle/ref_shim/t8mini.cpp (the t8code-API facade over this forest) */
int mf_face_neighbors(const mf_forest* f, int64_t e, int face, int64_t out[4]) { return face_neighbors(f, e, face, out); }
int mf_periodic(const mf_forest* f) { return f->periodic; }
void mf_element(const mf_forest* f, int64_t e, int* level, uint32_t coord[3]) {
  *level = f->level[e];
  unmorton(f->key[e], f->dim, coord);
}

/* do elements i .. i+2^d-1 form a family (same parent, all leaves of the same level, i is child 0)?
 * rank_offsets (may be NULL): a family split across ranks is not offered for coarsening. */
int mf_is_family(const mf_forest* f, int64_t i, const int64_t* rank_offsets, int nranks) {
  int dim = f->dim, nch = 1 << dim, l = f->level[i];
  if (l == 0 || i + nch > f->n) return 0;
  uint64_t pmask = ~(((uint64_t)1 << (dim * (MAXL - l + 1))) - 1);
  if (((f->key[i] >> (dim * (MAXL - l))) & (uint64_t)(nch - 1)) != 0) return 0;
  for (int c = 1; c < nch; c++)
    if (f->level[i + c] != l || (f->key[i + c] & pmask) != (f->key[i] & pmask)) return 0;
  if (rank_offsets)
    for (int p = 1; p < nranks; p++)
      if (rank_offsets[p] > i && rank_offsets[p] < i + nch) return 0;
  return 1;
}

/* apply per-element adapt results (1 refine, -1 coarsen the family starting here, 0 keep; entries of the other
 * family members are ignored when a family is coarsened), then 2:1 face balance (t8_forest_set_balance). */
mf_forest* mf_apply_adapt(const mf_forest* f, const int8_t* res) {
  int        dim = f->dim, nch = 1 << dim;
  mf_forest* g   = (mf_forest*)calloc(1, sizeof(mf_forest));
  g->dim         = dim;
  g->periodic    = f->periodic;
  g->key         = (uint64_t*)malloc(sizeof(uint64_t) * f->n * nch);
  g->level       = (int8_t*)malloc(f->n * nch);
  int64_t o      = 0;
  for (int64_t i = 0; i < f->n;) {
    int l = f->level[i];
    if (res[i] > 0) {
      for (int ch = 0; ch < nch; ch++) {
        g->key[o]   = f->key[i] | ((uint64_t)ch << (dim * (MAXL - l - 1)));
        g->level[o] = (int8_t)(l + 1);
        o++;
      }
      i += 1;
    } else if (res[i] < 0) {
      g->key[o]   = f->key[i];
      g->level[o] = (int8_t)(l - 1);
      o++;
      i += nch;
    } else {
      g->key[o]   = f->key[i];
      g->level[o] = (int8_t)l;
      o++;
      i += 1;
    }
  }
  g->n = o;
  return balance(g);
}

/* t8_forest_set_adapt(non-recursive) + set_balance, with the reference's callback (mesh_manager.inl:125-162 ==
 * subgrid_mesh_manager.inl:198-235): refine if level < max_level and crit > b; coarsen a family if level > min_level
 * and mean over the FIRST FOUR siblings (sic, also in 3-D) < b.  crit is float or double (is_f64). */
mf_forest* mf_adapt(const mf_forest* f, const void* crit, int is_f64, double b, int min_level, int max_level,
                    const int64_t* rank_offsets, int nranks) {
  int     nch = 1 << f->dim;
  int8_t* res = (int8_t*)calloc(f->n, 1);
  for (int64_t i = 0; i < f->n;) {
    int l = f->level[i];
    int is_family = mf_is_family(f, i, rank_offsets, nranks);
    int r = 0;
    if (l < max_level) {
      if (is_f64 ? (((const double*)crit)[i] > b) : (((const float*)crit)[i] > (float)b)) r = 1;
    }
    if (!r && l > min_level && is_family) {
      if (is_f64) {
        double c = 0.0;
        for (int s = 0; s < 4; s++) c += ((const double*)crit)[i + s] / 4.0;
        if (c < b) r = -1;
      } else {
        float c = 0.0f;
        for (int s = 0; s < 4; s++) c += ((const float*)crit)[i + s] / 4.0f;
        if (c < (float)b) r = -1;
      }
    }
    res[i] = (int8_t)r;
    i += r < 0 ? nch : 1;
  }
  mf_forest* g = mf_apply_adapt(f, res);
  free(res);
  return g;
}

/* mesh_manager.inl:258-281 (== subgrid_mesh_manager.inl:487-510): old->new element map from levels only.
 * adapt_data has n_new+1 entries. */
void mf_adapt_map(const mf_forest* fold, const mf_forest* fnew, int nb_subelements, int32_t* adapt_data) {
  int64_t oi = 0, ni = 0;
  while (oi < fold->n && ni < fnew->n) {
    int ol = fold->level[oi], nl = fnew->level[ni];
    if (ol < nl) {
      for (int i = 0; i < nb_subelements; i++) adapt_data[ni + i] = (int32_t)oi;
      oi += 1;
      ni += nb_subelements;
    } else if (ol > nl) {
      adapt_data[ni] = (int32_t)oi;
      oi += nb_subelements;
      ni += 1;
    } else {
      adapt_data[ni] = (int32_t)oi;
      oi += 1;
      ni += 1;
    }
  }
  adapt_data[ni] = (int32_t)oi;
}

void mf_partition_offsets(const mf_forest* f, int nranks, int64_t* offsets) {
  for (int p = 0; p <= nranks; p++) offsets[p] = (int64_t)(((__int128)f->n * p) / nranks);
}

/* ------------------------------------------------------------------ connectivity */
typedef struct {
  int64_t  n_local, n_ghost, n_faces, n_bfaces;
  int      ndim_normal;    /* components stored per normal */
  int32_t* ranks;          /* n_local + n_ghost */
  int32_t* indices;        /* n_local + n_ghost */
  int64_t* ghost_global;   /* n_ghost: global SFC index of each ghost (not part of the reference's arrays) */
  int32_t* face_neighbors; /* 2*n_faces + n_bfaces */
  double*  face_normals;   /* ndim_normal * (n_faces + n_bfaces), stored as double; exact for Cartesian */
  double*  face_areas;     /* n_faces + n_bfaces */
  int32_t* level_diff;     /* subgrid only: n_faces */
  int32_t* offsets;        /* subgrid only: dim * n_faces */
  /* extension (not in the reference): faces between a local element and a ghost owned by a LOWER rank, which the
   * reference leaves to that rank; needed by an owner-computes scheme.  Same record layout as the main list. */
  int64_t  n_xfaces;
  int32_t* x_face_neighbors;
  double*  x_face_normals;
  double*  x_face_areas;
  int32_t* x_level_diff;
  int32_t* x_offsets;
} mf_conn;

typedef struct { void* p; int64_t n, cap; size_t es; } vec;
static void vpush(vec* v, const void* x) {
  if (v->n == v->cap) {
    v->cap = v->cap ? 2 * v->cap : 1024;
    v->p   = realloc(v->p, v->cap * v->es);
  }
  memcpy((char*)v->p + v->n * v->es, x, v->es);
  v->n++;
}
static void vpush_i32(vec* v, int32_t x) { vpush(v, &x); }
static void vpush_f64(vec* v, double x) { vpush(v, &x); }

static int cmp_i64(const void* a, const void* b) {
  int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return x < y ? -1 : x > y;
}

static int owner_of(const int64_t* off, int nranks, int64_t g) {
  int lo = 0, hi = nranks - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= g)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

static void face_geom(const mf_forest* f, int64_t e, int face, double nrm[3], double* area) {
  double h = ldexp(1.0, -f->level[e]);
  nrm[0] = nrm[1] = nrm[2] = 0.0;
  nrm[face >> 1]           = (face & 1) ? 1.0 : -1.0;
  *area                    = f->dim == 3 ? h * h : h;
}

/* subgrid_mesh_manager.inl:560-786 add_face; E = subgrid extent (4). */
static void add_face_subgrid(const mf_forest* f, int face, int num_neighbors, int64_t eg, int32_t e_idx, int64_t ng,
                             int32_t n_idx, vec* ld, vec* off, vec* nbr, vec* nrmv, vec* areav) {
  const int E = 4;
  int       dim = f->dim, level = f->level[eg], nlevel = f->level[ng];
  double    nrm[3], area;
  face_geom(f, eg, face, nrm, &area);
  int level_difference = nlevel - level;
  int ax = face >> 1, o[3] = {0, 0, 0};
  if (level == nlevel) {
    o[ax] = (face & 1) ? 0 : E - 1;
    vpush_i32(ld, level_difference);
    for (int k = 0; k < dim; k++) vpush_i32(off, o[k]);
    vpush_i32(nbr, e_idx);
    vpush_i32(nbr, n_idx);
    for (int k = 0; k < dim; k++) vpush_f64(nrmv, nrm[k]);
  } else if (nlevel < level) {
    int child_id = (int)((f->key[eg] >> (dim * (MAXL - level))) & ((1u << dim) - 1));
    for (int d = 0; d < dim; d++) o[d] = E / 2 * ((child_id >> d) & 1);
    o[ax] = (face & 1) ? 0 : E - 1;
    vpush_i32(ld, level_difference);
    for (int k = 0; k < dim; k++) vpush_i32(off, o[k]);
    vpush_i32(nbr, e_idx);
    vpush_i32(nbr, n_idx);
    for (int k = 0; k < dim; k++) vpush_f64(nrmv, nrm[k]);
  } else {
    int child_id = (int)((f->key[ng] >> (dim * (MAXL - nlevel))) & ((1u << dim) - 1));
    for (int d = 0; d < dim; d++) o[d] = E / 2 * ((child_id >> d) & 1);
    o[ax] = (face & 1) ? E - 1 : 0; /* (face_idx == 0|2|4) ? 0 : E-1 */
    vpush_i32(ld, -level_difference);
    for (int k = 0; k < dim; k++) vpush_i32(off, o[k]);
    vpush_i32(nbr, n_idx);
    vpush_i32(nbr, e_idx);
    for (int k = 0; k < dim; k++) vpush_f64(nrmv, -nrm[k]);
  }
  vpush_f64(areav, area / (double)num_neighbors);
}

/* subgrid = 0: MeshManager<.., dim_normal> (mesh_manager.inl:332-481); subgrid = 1: SubgridMeshManager.
 * NOTE on area/num_neighbors: the reference divides two float_type values; for Cartesian forests both operands are
 * powers of two / small integers so the double result cast afterwards is identical. */
mf_conn* mf_connectivity(const mf_forest* f, int nranks, int rank, int subgrid, int ndim_normal) {
  int      dim = f->dim;
  int64_t* off = (int64_t*)malloc(sizeof(int64_t) * (nranks + 1));
  mf_partition_offsets(f, nranks, off);
  int64_t lo = off[rank], hi = off[rank + 1], nl = hi - lo;

  /* ghost layer */
  vec gh = {0, 0, 0, sizeof(int64_t)};
  for (int64_t e = lo; e < hi; e++)
    for (int face = 0; face < 2 * dim; face++) {
      int64_t nb[4];
      int     nn = face_neighbors(f, e, face, nb);
      for (int i = 0; i < nn; i++)
        if (nb[i] < lo || nb[i] >= hi) vpush(&gh, &nb[i]);
    }
  qsort(gh.p, gh.n, sizeof(int64_t), cmp_i64);
  int64_t ng = 0;
  for (int64_t i = 0; i < gh.n; i++)
    if (i == 0 || ((int64_t*)gh.p)[i] != ((int64_t*)gh.p)[i - 1]) ((int64_t*)gh.p)[ng++] = ((int64_t*)gh.p)[i];
  int64_t* ghosts = (int64_t*)gh.p;

  mf_conn* c      = (mf_conn*)calloc(1, sizeof(mf_conn));
  c->n_local      = nl;
  c->n_ghost      = ng;
  c->ndim_normal  = subgrid ? dim : ndim_normal;
  c->ranks        = (int32_t*)malloc(sizeof(int32_t) * (nl + ng + 1));
  c->indices      = (int32_t*)malloc(sizeof(int32_t) * (nl + ng + 1));
  c->ghost_global = (int64_t*)malloc(sizeof(int64_t) * (ng + 1));
  for (int64_t i = 0; i < nl; i++) {
    c->ranks[i]   = rank;
    c->indices[i] = (int32_t)i;
  }
  for (int64_t g = 0; g < ng; g++) {
    int p              = owner_of(off, nranks, ghosts[g]);
    c->ranks[nl + g]   = p;
    c->indices[nl + g] = (int32_t)(ghosts[g] - off[p]);
    c->ghost_global[g] = ghosts[g];
  }

  vec nbr = {0, 0, 0, 4}, bnbr = {0, 0, 0, 4}, nrm = {0, 0, 0, 8}, bnrm = {0, 0, 0, 8}, ar = {0, 0, 0, 8},
      bar = {0, 0, 0, 8}, ld = {0, 0, 0, 4}, ofs = {0, 0, 0, 4};
  vec xnbr = {0, 0, 0, 4}, xnrm = {0, 0, 0, 8}, xar = {0, 0, 0, 8}, xld = {0, 0, 0, 4}, xofs = {0, 0, 0, 4};
  int nd = c->ndim_normal;

  for (int64_t e = lo; e < hi; e++) {
    int32_t e_idx = (int32_t)(e - lo);
    for (int face = 0; face < 2 * dim; face++) {
      int64_t nb[4];
      int32_t nid[4];
      int     nn = face_neighbors(f, e, face, nb);
      double  nv[3], area;
      face_geom(f, e, face, nv, &area);
      for (int i = 0; i < nn; i++) {
        if (nb[i] >= lo && nb[i] < hi)
          nid[i] = (int32_t)(nb[i] - lo);
        else {
          int64_t* p = (int64_t*)bsearch(&nb[i], ghosts, ng, sizeof(int64_t), cmp_i64);
          nid[i]     = (int32_t)(nl + (p - ghosts));
        }
      }
      /* ghost neighbours first (mesh_manager.inl:396-409 / subgrid_mesh_manager.inl:859-877) */
      for (int i = 0; i < nn; i++) {
        if (nid[i] >= nl) {
          int own = rank < c->ranks[nid[i]];
          if (subgrid) {
            if (own)
              add_face_subgrid(f, face, nn, e, e_idx, nb[i], nid[i], &ld, &ofs, &nbr, &nrm, &ar);
            else
              add_face_subgrid(f, face, nn, e, e_idx, nb[i], nid[i], &xld, &xofs, &xnbr, &xnrm, &xar);
          } else {
            vec *pn = own ? &nbr : &xnbr, *pr = own ? &nrm : &xnrm, *pa = own ? &ar : &xar;
            vpush_i32(pn, e_idx);
            vpush_i32(pn, nid[i]);
            for (int k = 0; k < nd; k++) vpush_f64(pr, nv[k]);
            vpush_f64(pa, area / (double)nn);
          }
        }
      }
      /* local neighbour (mesh_manager.inl:411-424) */
      if (nn == 1 && nid[0] < nl &&
          (nid[0] > e_idx || (nid[0] < e_idx && f->level[nb[0]] < f->level[e]))) {
        if (subgrid) {
          add_face_subgrid(f, face, nn, e, e_idx, nb[0], nid[0], &ld, &ofs, &nbr, &nrm, &ar);
        } else {
          vpush_i32(&nbr, e_idx);
          vpush_i32(&nbr, nid[0]);
          for (int k = 0; k < nd; k++) vpush_f64(&nrm, nv[k]);
          vpush_f64(&ar, area);
        }
      }
      if (nn == 0) {
        vpush_i32(&bnbr, e_idx);
        for (int k = 0; k < nd; k++) vpush_f64(&bnrm, nv[k]);
        vpush_f64(&bar, area);
      }
    }
  }
  c->n_faces  = ar.n;
  c->n_bfaces = bar.n;
  c->face_neighbors = (int32_t*)malloc(4 * (nbr.n + bnbr.n + 1));
  memcpy(c->face_neighbors, nbr.p, 4 * nbr.n);
  memcpy(c->face_neighbors + nbr.n, bnbr.p, 4 * bnbr.n);
  c->face_normals = (double*)malloc(8 * (nrm.n + bnrm.n + 1));
  memcpy(c->face_normals, nrm.p, 8 * nrm.n);
  memcpy(c->face_normals + nrm.n, bnrm.p, 8 * bnrm.n);
  c->face_areas = (double*)malloc(8 * (ar.n + bar.n + 1));
  memcpy(c->face_areas, ar.p, 8 * ar.n);
  memcpy(c->face_areas + ar.n, bar.p, 8 * bar.n);
  c->level_diff = (int32_t*)ld.p;
  c->offsets    = (int32_t*)ofs.p;
  c->n_xfaces         = xar.n;
  c->x_face_neighbors = (int32_t*)xnbr.p;
  c->x_face_normals   = (double*)xnrm.p;
  c->x_face_areas     = (double*)xar.p;
  c->x_level_diff     = (int32_t*)xld.p;
  c->x_offsets        = (int32_t*)xofs.p;
  free(nbr.p);
  free(bnbr.p);
  free(nrm.p);
  free(bnrm.p);
  free(ar.p);
  free(bar.p);
  free(gh.p);
  free(off);
  return c;
}

void mf_conn_free(mf_conn* c) {
  if (!c) return;
  free(c->ranks);
  free(c->indices);
  free(c->ghost_global);
  free(c->face_neighbors);
  free(c->face_normals);
  free(c->face_areas);
  free(c->level_diff);
  free(c->offsets);
  free(c->x_face_neighbors);
  free(c->x_face_normals);
  free(c->x_face_areas);
  free(c->x_level_diff);
  free(c->x_offsets);
  free(c);
}
