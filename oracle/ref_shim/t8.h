/* TEST INFRASTRUCTURE ONLY -- "t8mini": declarations of the subset of the t8code / libsc / MPI API that the t8gpu
 * reference calls, so that the reference's translation units compile UNMODIFIED from
 * /root/reference.  Definitions live in t8mini.cpp (single process, one rank, Cartesian one-tree forests built on
 * oracle/miniforest.c).  t8code, libsc and MPI are not installed in this image; this is NOT t8code. */
#ifndef T8MINI_T8_H
#define T8MINI_T8_H
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

/* ---- MPI (single process) */
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
#define MPI_COMM_WORLD 0
#define MPI_IN_PLACE ((void*)1)
#define MPI_DATATYPE_NULL 0
#define MPI_BYTE 1
#define MPI_FLOAT 2
#define MPI_DOUBLE 3
#define MPI_INT 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_SUCCESS 0
#ifdef __cplusplus
extern "C" {
#endif
int MPI_Comm_size(MPI_Comm, int*);
int MPI_Comm_rank(MPI_Comm, int*);
int MPI_Barrier(MPI_Comm);
int MPI_Allgather(const void*, int, MPI_Datatype, void*, int, MPI_Datatype, MPI_Comm);
int MPI_Allreduce(const void*, void*, int, MPI_Datatype, MPI_Op, MPI_Comm);

/* ---- libsc */
typedef MPI_Comm sc_MPI_Comm;
#define sc_MPI_COMM_WORLD MPI_COMM_WORLD
typedef struct sc_array { void* array; size_t elem_size, elem_count; } sc_array;
sc_array* sc_array_new_data(void* base, size_t elem_size, size_t elem_count);
void      sc_array_destroy(sc_array*);
#define SC_ABORT(msg) do { fprintf(stderr, "SC_ABORT: %s\n", msg); abort(); } while (0)
#define SC_CHECK_MPI(r) do { if ((r) != MPI_SUCCESS) SC_ABORT("MPI error"); } while (0)
#define SC_LP_ESSENTIAL 7
#define SC_LP_PRODUCTION 6
int  sc_MPI_Init(int*, char***);
int  sc_MPI_Finalize(void);
void sc_init(sc_MPI_Comm, int, int, void*, int);
void sc_finalize(void);
void t8_init(int);

/* ---- t8code */
typedef int32_t t8_locidx_t;
typedef int64_t t8_gloidx_t;
typedef struct t8mini_forest* t8_forest_t;
typedef struct t8mini_cmesh*  t8_cmesh_t;
typedef struct t8mini_element t8_element_t;
typedef struct t8mini_scheme  t8_scheme_cxx_t;
typedef enum { T8_ECLASS_QUAD = 3, T8_ECLASS_HEX = 4 } t8_eclass_t;
typedef enum { T8_GHOST_NONE = 0, T8_GHOST_FACES } t8_ghost_type_t;
typedef enum { T8_VTK_SCALAR, T8_VTK_VECTOR } t8_vtk_data_type_t;
typedef struct { t8_vtk_data_type_t type; char description[BUFSIZ]; double* data; } t8_vtk_data_field_t;
#define T8_FREE(p) free(p)
#ifdef __cplusplus
}
struct t8_eclass_scheme_c {
  int dim;
  int  t8_element_level(const t8_element_t* e);
  int  t8_element_num_faces(const t8_element_t* e);
  int  t8_element_child_id(const t8_element_t* e);
  void t8_element_destroy(int n, t8_element_t** e);
};
typedef int (*t8_forest_adapt_t)(t8_forest_t forest, t8_forest_t forest_from, t8_locidx_t which_tree,
                                 t8_locidx_t lelement_id, t8_eclass_scheme_c* ts, const int is_family,
                                 const int num_elements, t8_element_t* elements[]);
extern "C" {
t8_scheme_cxx_t* t8_scheme_new_default_cxx(void);
t8_cmesh_t  t8_cmesh_new_periodic(sc_MPI_Comm comm, int dim);
t8_cmesh_t  t8_cmesh_new_prismed_spherical_shell_icosahedron(double, double, int, int, sc_MPI_Comm);
void        t8_cmesh_destroy(t8_cmesh_t* pcmesh);
t8_forest_t t8_forest_new_uniform(t8_cmesh_t cmesh, t8_scheme_cxx_t* scheme, int level, int do_face_ghost,
                                  sc_MPI_Comm comm);
void  t8_forest_init(t8_forest_t*);
void  t8_forest_ref(t8_forest_t);
void  t8_forest_unref(t8_forest_t*);
void  t8_forest_commit(t8_forest_t);
int   t8_forest_is_committed(t8_forest_t);
void  t8_forest_set_adapt(t8_forest_t forest, t8_forest_t set_from, t8_forest_adapt_t fn, int recursive);
void  t8_forest_set_ghost(t8_forest_t forest, int do_ghost, t8_ghost_type_t type);
void  t8_forest_set_balance(t8_forest_t forest, t8_forest_t set_from, int no_repartition);
void  t8_forest_set_partition(t8_forest_t forest, t8_forest_t set_from, int set_for_coarsening);
void  t8_forest_set_user_data(t8_forest_t, void*);
void* t8_forest_get_user_data(t8_forest_t);
t8_locidx_t t8_forest_get_local_num_elements(t8_forest_t);
t8_locidx_t t8_forest_get_num_ghosts(t8_forest_t);
t8_locidx_t t8_forest_get_num_local_trees(t8_forest_t);
t8_eclass_t t8_forest_get_tree_class(t8_forest_t, t8_locidx_t ltreeid);
t8_eclass_scheme_c* t8_forest_get_eclass_scheme(t8_forest_t, t8_eclass_t);
t8_locidx_t t8_forest_get_tree_num_elements(t8_forest_t, t8_locidx_t ltreeid);
t8_locidx_t t8_forest_get_tree_element_offset(t8_forest_t, t8_locidx_t ltreeid);
t8_element_t* t8_forest_get_element_in_tree(t8_forest_t, t8_locidx_t ltreeid, t8_locidx_t leid_in_tree);
double t8_forest_element_volume(t8_forest_t, t8_locidx_t ltreeid, const t8_element_t*);
double t8_forest_element_face_area(t8_forest_t, t8_locidx_t ltreeid, const t8_element_t*, int face);
void   t8_forest_element_face_normal(t8_forest_t, t8_locidx_t ltreeid, const t8_element_t*, int face, double normal[3]);
void   t8_forest_element_centroid(t8_forest_t, t8_locidx_t ltreeid, const t8_element_t*, double* coordinates);
void   t8_forest_leaf_face_neighbors(t8_forest_t, t8_locidx_t ltreeid, const t8_element_t* leaf,
                                     t8_element_t** pneighbor_leafs[], int face, int* dual_faces[], int* num_neighbors,
                                     t8_locidx_t** pelement_indices, t8_eclass_scheme_c** pneigh_scheme,
                                     int forest_is_balanced);
void   t8_forest_ghost_exchange_data(t8_forest_t, sc_array* element_data);
void   t8_forest_partition_data(t8_forest_t forest_from, t8_forest_t forest_to, const sc_array* data_in,
                                sc_array* data_out);
int    t8_forest_write_vtk_ext(t8_forest_t, const char* prefix, int write_treeid, int write_mpirank, int write_level,
                               int write_element_id, int write_ghosts, int write_curved, int do_not_use_API,
                               int num_data, t8_vtk_data_field_t* data);
/* t8mini extension used by the harness only */
t8_cmesh_t t8mini_cmesh_new_cube(int dim, int periodic);
}
#endif
#endif
