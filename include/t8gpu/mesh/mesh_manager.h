/// @file mesh_manager.h
/// @brief MeshManager and MeshConnectivityAccessor: source-compatible with t8gpu/mesh/mesh_manager.h:29-465 (same
///        template parameters, member names, return types usable in `__global__` signatures).
///
/// What is the same as in the reference (and has to be, for its connectivity arrays to come out bit for bit): the t8code
/// calls that own the forest (adapt with the same callback, balance, partition, face-ghost layer) and the enumeration
/// order of faces -- elements in SFC order x face id, ghost-neighbour faces before the local-neighbour face, hanging
/// faces emitted from the fine side (t8gpu/mesh/mesh_manager.inl:332-481).
///
/// What is different (behind the same interface):
///  - one generic leaf walk (`for_each_leaf`) serves initialisation, the level walk of adapt() and the connectivity;
///  - compute_connectivity_information() also records the partition-boundary faces whose ghost neighbour is owned by a
///    LOWER rank (the reference leaves those to that rank, which then writes this rank's fluxes with remote atomics)
///    and builds the tile plan of the fused stage kernel (t8b200_plan_create): every rank evaluates all faces of its own
///    elements, no remote atomics;
///  - adapt() / partition() remap variables and volumes on the device straight into the new allocation
///    (t8b200_adapt_remap / t8b200_partition_remap) instead of a temporary buffer + one device-to-device copy per
///    variable, and the new allocation is zero-initialised;
///  - fused_stage() / gradient_criteria() expose the fused kernels; the reference's kernels still run on the accessors.
#ifndef T8GPU_B200_MESH_MESH_MANAGER_H
#define T8GPU_B200_MESH_MESH_MANAGER_H

#include <t8.h>
#include <t8_cmesh.h>
#include <t8_forest/t8_forest.h>
#include <t8_forest/t8_forest_io.h>
#include <t8_forest/t8_forest_partition.h>
#include <t8gpu/memory/memory_manager.h>
#include <t8gpu/utils/cuda.h>
#include <t8gpu_b200.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <array>
#include <cassert>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace t8gpu {

  /// Kernel-argument POD with the face connectivity of one rank (mesh_manager.h:29-182).
  template<typename float_type, size_t dim>
  class MeshConnectivityAccessor {
    template<typename VT, typename ST, size_t dim_>
    friend class MeshManager;

   public:
    MeshConnectivityAccessor(MeshConnectivityAccessor const&)            = default;
    MeshConnectivityAccessor& operator=(MeshConnectivityAccessor const&) = default;

    [[nodiscard]] __device__ __host__ inline t8_locidx_t get_num_local_faces() const { return m_num_local_faces; }
    [[nodiscard]] __device__ __host__ inline t8_locidx_t get_num_local_boundary_faces() const {
      return m_num_local_boundary_faces;
    }
    [[nodiscard]] __device__ inline float_type get_face_surface(int face_idx) const { return m_face_surfaces[face_idx]; }
    [[nodiscard]] __device__ inline float_type get_boundary_face_surface(int face_idx) const {
      return m_face_surfaces[m_num_local_faces + face_idx];
    }
    [[nodiscard]] __device__ inline std::array<float_type, dim> get_face_normal(int face_idx) const {
      std::array<float_type, dim> n{};
      for (size_t k = 0; k < dim; k++) n[k] = m_face_normals[dim * face_idx + k];
      return n;
    }
    [[nodiscard]] __device__ inline std::array<float_type, dim> get_boundary_face_normal(int face_idx) const {
      return get_face_normal(m_num_local_faces + face_idx);
    }
    [[nodiscard]] __device__ inline std::array<t8_locidx_t, 2> get_face_neighbor_indices(int face_idx) const {
      return {m_face_neighbors[2 * face_idx], m_face_neighbors[2 * face_idx + 1]};
    }
    [[nodiscard]] __device__ inline t8_locidx_t get_boundary_face_neighbor_index(int face_idx) const {
      return m_face_neighbors[2 * m_num_local_faces + face_idx];
    }
    [[nodiscard]] __device__ inline t8_locidx_t get_element_owner_rank(int element_idx) const {
      return m_ranks[element_idx];
    }
    [[nodiscard]] __device__ inline t8_locidx_t get_element_owner_remote_index(int element_idx) const {
      return m_indices[element_idx];
    }

    /// raw device arrays, as the C ABI (t8gpu_b200.h) takes them (t8gpu_b200 extension)
    [[nodiscard]] __host__ int const*         ranks() const { return m_ranks; }
    [[nodiscard]] __host__ t8_locidx_t const* indices() const { return m_indices; }
    [[nodiscard]] __host__ t8_locidx_t const* face_neighbors() const { return m_face_neighbors; }
    [[nodiscard]] __host__ float_type const*  face_normals() const { return m_face_normals; }
    [[nodiscard]] __host__ float_type const*  face_surfaces() const { return m_face_surfaces; }

   private:
    int const*         m_ranks;
    t8_locidx_t const* m_indices;
    t8_locidx_t const* m_face_neighbors;
    float_type const*  m_face_normals;
    float_type const*  m_face_surfaces;
    t8_locidx_t        m_num_local_faces;
    t8_locidx_t        m_num_local_boundary_faces;

    MeshConnectivityAccessor(int const* ranks, t8_locidx_t const* indices, t8_locidx_t const* face_neighbors,
                             float_type const* face_normals, float_type const* face_surfaces, t8_locidx_t nf,
                             t8_locidx_t nb)
        : m_ranks{ranks}, m_indices{indices}, m_face_neighbors{face_neighbors}, m_face_normals{face_normals},
          m_face_surfaces{face_surfaces}, m_num_local_faces{nf}, m_num_local_boundary_faces{nb} {}
  };

  /// Owns a t8code forest and the device data attached to its elements (mesh_manager.h:231-465).
  template<typename VariableType, typename StepType, size_t dim>
  class MeshManager : public MemoryManager<VariableType, StepType> {
    using Memory = MemoryManager<VariableType, StepType>;

   public:
    using float_type                  = typename variable_traits<VariableType>::float_type;
    using variable_index_type         = typename variable_traits<VariableType>::index_type;
    static constexpr int nb_variables = variable_traits<VariableType>::nb_variables;
    using step_index_type             = typename step_traits<StepType>::index_type;
    static constexpr size_t nb_steps  = step_traits<StepType>::nb_steps;

    static constexpr t8_locidx_t min_level = 1;
    static constexpr t8_locidx_t max_level = 4;

    /// Takes ownership of cmesh and forest (freed in the destructor); collective.
    MeshManager(sc_MPI_Comm comm, t8_scheme_cxx_t* scheme, t8_cmesh_t cmesh, t8_forest_t forest)
        : Memory{static_cast<size_t>(t8_forest_get_local_num_elements(forest)), comm},
          m_comm{comm}, m_scheme{scheme}, m_cmesh{cmesh}, m_forest{forest} {
      MPI_Comm_size(m_comm, &m_nb_ranks);
      MPI_Comm_rank(m_comm, &m_rank);
      refresh_counts();
      m_user_data.element_refinement_criteria = &m_element_refinement_criteria;
      t8_forest_set_user_data(m_forest, &m_user_data);
      m_element_refinement_criteria.resize(m_num_local_elements);
      compute_connectivity_information();
    }
    ~MeshManager() {
      t8b200_plan_destroy(m_plan);
      t8_forest_unref(&m_forest);
      t8_cmesh_destroy(&m_cmesh);
    }
    MeshManager(MeshManager const&)            = delete;
    MeshManager& operator=(MeshManager const&) = delete;

    /// func(MemoryAccessorOwn& host_variables, t8_forest_t, t8_locidx_t tree, t8_element_t const*, t8_locidx_t index)
    /// fills the variables of step 0 on the host; volumes come from t8code (mesh_manager.h:266-276).
    template<typename Func>
    void initialize_variables(Func func) {
      std::array<thrust::host_vector<float_type>, nb_variables> host{};
      std::array<float_type*, nb_variables>                     rows{};
      for (int k = 0; k < nb_variables; k++) {
        host[k].resize(m_num_local_elements);
        rows[k] = host[k].data();
      }
      MemoryAccessorOwn<VariableType> host_accessor{rows};
      thrust::host_vector<float_type> volume(m_num_local_elements);
      for_each_leaf(m_forest, [&](t8_locidx_t tree, t8_eclass_scheme_c*, t8_element_t const* element, t8_locidx_t idx) {
        volume[idx] = static_cast<float_type>(t8_forest_element_volume(m_forest, tree, element));
        func(host_accessor, m_forest, tree, element, idx);
      });
      Memory fresh{static_cast<size_t>(m_num_local_elements), m_comm};   // zero-initialised
      static_cast<Memory&>(*this) = std::move(fresh);
      for (int k = 0; k < nb_variables; k++)
        this->set_variable(static_cast<step_index_type>(0), static_cast<variable_index_type>(k), host[k]);
      this->set_volume(volume);
    }

    /// t8code adapt (+ 2:1 balance, face ghosts) driven by one criterion per element, then device remap of the
    /// variables of `step` and of the volumes.  Collective.
    void adapt(thrust::host_vector<float_type> const& refinement_criteria, step_index_type step) {
      assert(t8_forest_is_committed(m_forest));
      assert(static_cast<t8_locidx_t>(refinement_criteria.size()) == m_num_local_elements);
      m_element_refinement_criteria = refinement_criteria;

      t8_forest_ref(m_forest);
      t8_forest_t adapted{};
      t8_forest_init(&adapted);
      t8_forest_set_adapt(adapted, m_forest, adapt_callback_iteration, false);
      t8_forest_set_ghost(adapted, true, T8_GHOST_FACES);
      t8_forest_set_balance(adapted, m_forest, true);
      t8_forest_commit(adapted);

      // old -> new index map from the element levels (a family is nb_children consecutive leaves)
      std::vector<int> const old_levels = leaf_levels(m_forest), new_levels = leaf_levels(adapted);
      constexpr int          nb_children = dim == 2 ? 4 : 8;
      t8_locidx_t const      n_old = static_cast<t8_locidx_t>(old_levels.size()),
                        n_new      = static_cast<t8_locidx_t>(new_levels.size());
      thrust::host_vector<t8_locidx_t> map(n_new + 1);
      t8_locidx_t                      o = 0, n = 0;
      while (o < n_old && n < n_new) {
        if (old_levels[o] < new_levels[n]) {          // refined: the children all point at their parent
          for (int c = 0; c < nb_children; c++) map[n + c] = o;
          o += 1;
          n += nb_children;
        } else if (old_levels[o] > new_levels[n]) {   // coarsened: the parent points at its first child
          map[n] = o;
          o += nb_children;
          n += 1;
        } else {
          map[n++] = o++;
        }
      }
      map[n] = o;

      thrust::device_vector<t8_locidx_t> device_map = map;
      Memory fresh{static_cast<size_t>(n_new), m_comm};
      T8GPU_CUDA_CHECK_ERROR(remap_adapt(thrust::raw_pointer_cast(device_map.data()), n_new, step, fresh));
      T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize());
      MPI_Barrier(m_comm);
      static_cast<Memory&>(*this) = std::move(fresh);
      m_element_refinement_criteria.resize(n_new);

      t8_forest_set_user_data(adapted, &m_user_data);
      t8_forest_unref(&m_forest);
      m_forest = adapted;
      refresh_counts();
    }

    /// t8code repartition, then every rank pulls its new elements from the ranks that held them.  Collective.
    void partition(step_index_type step) {
      assert(t8_forest_is_committed(m_forest));
      t8_forest_ref(m_forest);
      t8_forest_t partitioned{};
      t8_forest_init(&partitioned);
      t8_forest_set_partition(partitioned, m_forest, true);
      t8_forest_set_ghost(partitioned, true, T8_GHOST_FACES);
      t8_forest_commit(partitioned);

      t8_locidx_t const n_old = t8_forest_get_local_num_elements(m_forest),
                        n_new = t8_forest_get_local_num_elements(partitioned);
      thrust::host_vector<int>         old_ranks(n_old, m_rank), new_ranks(n_new);
      thrust::host_vector<t8_locidx_t> old_indices(n_old), new_indices(n_new);
      for (t8_locidx_t i = 0; i < n_old; i++) old_indices[i] = i;
      ship(m_forest, partitioned, old_ranks, new_ranks);
      ship(m_forest, partitioned, old_indices, new_indices);

      thrust::device_vector<int>         d_ranks   = new_ranks;
      thrust::device_vector<t8_locidx_t> d_indices = new_indices;
      Memory fresh{static_cast<size_t>(n_new), m_comm};
      T8GPU_CUDA_CHECK_ERROR(t8b200_partition_remap(
          nb_variables, n_new, 1, thrust::raw_pointer_cast(d_ranks.data()), thrust::raw_pointer_cast(d_indices.data()),
          fresh.get_own_variables(step).data(), this->get_all_variables(step).data(), fresh.get_own_volume(),
          this->get_all_volume(), nullptr));
      T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize());
      MPI_Barrier(m_comm);   // nobody frees its old allocation while a peer still reads it
      static_cast<Memory&>(*this) = std::move(fresh);
      m_element_refinement_criteria.resize(n_new);

      t8_forest_set_user_data(partitioned, &m_user_data);
      t8_forest_unref(&m_forest);
      m_forest = partitioned;
      refresh_counts();
    }

    /// Rebuilds the owner tables, the face arrays (bit for bit the reference's) and the tile plan.  Collective.
    void compute_connectivity_information() {
      assert(t8_forest_is_committed(m_forest));
      t8_locidx_t const n_all = m_num_local_elements + m_num_ghost_elements;
      m_ranks.assign(n_all, m_rank);
      m_indices.resize(n_all);
      for (t8_locidx_t i = 0; i < m_num_local_elements; i++) m_indices[i] = i;
      ghost_exchange(m_ranks);
      ghost_exchange(m_indices);
      m_device_ranks   = m_ranks;
      m_device_indices = m_indices;

      // interior faces, boundary faces, and the partition-boundary faces owned by a lower rank ("x")
      std::vector<t8_locidx_t> nbr, bnbr, xnbr;
      std::vector<float_type>  nrm, bnrm, xnrm, area, barea, xarea;
      auto emit = [&](std::vector<t8_locidx_t>& ids, std::vector<float_type>& normals, std::vector<float_type>& areas,
                      t8_locidx_t tree, t8_element_t const* element, int face, t8_locidx_t left, t8_locidx_t right,
                      int share) {
        ids.push_back(left);
        if (right >= 0) ids.push_back(right);
        double n[3] = {0.0, 0.0, 0.0};
        t8_forest_element_face_normal(m_forest, tree, element, face, n);
        for (size_t k = 0; k < dim; k++) normals.push_back(static_cast<float_type>(n[k]));
        float_type a = static_cast<float_type>(t8_forest_element_face_area(m_forest, tree, element, face));
        areas.push_back(share > 1 ? a / static_cast<float_type>(share) : a);
      };
      for_each_leaf(m_forest, [&](t8_locidx_t tree, t8_eclass_scheme_c* scheme, t8_element_t const* element,
                                  t8_locidx_t idx) {
        int const num_faces = scheme->t8_element_num_faces(element);
        for (int face = 0; face < num_faces; face++) {
          int                 num_neighbors = 0;
          int*                dual_faces    = nullptr;
          t8_locidx_t*        ids           = nullptr;
          t8_element_t**      neighbors     = nullptr;
          t8_eclass_scheme_c* nscheme       = nullptr;
          t8_forest_leaf_face_neighbors(m_forest, tree, element, &neighbors, face, &dual_faces, &num_neighbors, &ids,
                                        &nscheme, true);
          for (int i = 0; i < num_neighbors; i++) {   // faces to ghosts: owned by the lower rank, mirrored as "x"
            if (ids[i] < m_num_local_elements) continue;
            if (m_rank < m_ranks[ids[i]]) emit(nbr, nrm, area, tree, element, face, idx, ids[i], num_neighbors);
            else emit(xnbr, xnrm, xarea, tree, element, face, idx, ids[i], num_neighbors);
          }
          if (num_neighbors == 1 && ids[0] < m_num_local_elements) {
            // conforming faces once (from the lower index), hanging faces from the fine side
            bool const lower = ids[0] > idx;
            bool const finer = ids[0] < idx && nscheme->t8_element_level(neighbors[0]) < scheme->t8_element_level(element);
            if (lower || finer) emit(nbr, nrm, area, tree, element, face, idx, ids[0], 1);
          }
          if (num_neighbors == 0) emit(bnbr, bnrm, barea, tree, element, face, idx, -1, 1);
          if (neighbors) {
            nscheme->t8_element_destroy(num_neighbors, neighbors);
            T8_FREE(neighbors);
          }
          T8_FREE(dual_faces);
          T8_FREE(ids);
        }
      });
      m_num_local_faces          = static_cast<t8_locidx_t>(area.size());
      m_num_local_boundary_faces = static_cast<t8_locidx_t>(barea.size());
      m_num_x_faces              = static_cast<t8_locidx_t>(xarea.size());
      nbr.insert(nbr.end(), bnbr.begin(), bnbr.end());
      nrm.insert(nrm.end(), bnrm.begin(), bnrm.end());
      area.insert(area.end(), barea.begin(), barea.end());
      m_device_face_neighbors.assign(nbr.begin(), nbr.end());
      m_device_face_normals.assign(nrm.begin(), nrm.end());
      m_device_face_area.assign(area.begin(), area.end());

      // tile plan of the fused stage kernel (normals with 3 components)
      auto to3 = [](std::vector<float_type> const& v) {
        if (dim == 3) return v;
        std::vector<float_type> out(v.size() / dim * 3, float_type(0));
        for (size_t f = 0; f < v.size() / dim; f++)
          for (size_t k = 0; k < dim; k++) out[3 * f + k] = v[dim * f + k];
        return out;
      };
      std::vector<float_type> const nrm3 = to3(nrm), xnrm3 = to3(xnrm);
      t8b200_plan_destroy(m_plan);
      m_plan = nullptr;
      T8GPU_CUDA_CHECK_ERROR(t8b200_plan_create(
          &m_plan, sizeof(float_type) == 8, m_num_local_elements, m_num_ghost_elements, m_num_local_faces,
          m_num_local_boundary_faces, nbr.data(), nrm3.data(), area.data(), m_ranks.data(), m_indices.data(),
          m_num_x_faces, xnbr.data(), xnrm3.data(), xarea.data()));
    }

    // ---- output (mesh_manager.h:305-372)
    class HostVariableInfo {
      friend MeshManager;
      HostVariableInfo(t8_vtk_data_type_t type, std::unique_ptr<double[]>&& data, std::string const& name)
          : m_data{std::move(data)} {
        m_field.type = type;
        m_field.data = m_data.get();
        std::strncpy(m_field.description, name.c_str(), BUFSIZ - 1);
        m_field.description[BUFSIZ - 1] = '\0';
      }

     public:
      HostVariableInfo()                   = default;
      HostVariableInfo(HostVariableInfo&&) = default;
      ~HostVariableInfo()                  = default;

     private:
      std::unique_ptr<double[]> m_data;
      t8_vtk_data_field_t       m_field;
    };

    [[nodiscard]] HostVariableInfo get_host_scalar_variable(step_index_type step, variable_index_type variable,
                                                            std::string const& name) const {
      auto data = std::make_unique<double[]>(m_num_local_elements);
      fetch(step, variable, data.get(), 1);
      return HostVariableInfo{T8_VTK_SCALAR, std::move(data), name};
    }
    [[nodiscard]] HostVariableInfo get_host_vector_variable(step_index_type step, std::array<variable_index_type, 3> variable,
                                                            std::string const& name) const {
      auto data = std::make_unique<double[]>(3 * static_cast<size_t>(m_num_local_elements));
      for (int c = 0; c < 3; c++) fetch(step, variable[c], data.get() + c, 3);
      return HostVariableInfo{T8_VTK_VECTOR, std::move(data), name};
    }
    void save_variables_to_vtk(std::vector<HostVariableInfo> host_variables, std::string const& prefix) const {
      std::vector<t8_vtk_data_field_t> fields(host_variables.size());
      for (size_t k = 0; k < host_variables.size(); k++) fields[k] = host_variables[k].m_field;
      t8_forest_write_vtk_ext(m_forest, prefix.c_str(), true, true, true, true, false, false, false,
                              static_cast<int>(fields.size()), fields.data());
    }
    void save_variable_to_vtk(step_index_type step, variable_index_type variable, std::string const& prefix) const {
      std::vector<HostVariableInfo> v;
      v.push_back(get_host_scalar_variable(step, variable, "variable"));
      save_variables_to_vtk(std::move(v), prefix);
    }

    [[nodiscard]] MeshConnectivityAccessor<float_type, dim> get_connectivity_information() const {
      return {thrust::raw_pointer_cast(m_device_ranks.data()), thrust::raw_pointer_cast(m_device_indices.data()),
              thrust::raw_pointer_cast(m_device_face_neighbors.data()), thrust::raw_pointer_cast(m_device_face_normals.data()),
              thrust::raw_pointer_cast(m_device_face_area.data()), m_num_local_faces, m_num_local_boundary_faces};
    }
    [[nodiscard]] t8_locidx_t get_num_local_elements() const { return m_num_local_elements; }
    [[nodiscard]] t8_locidx_t get_num_ghost_elements() const { return m_num_ghost_elements; }
    [[nodiscard]] t8_locidx_t get_num_local_faces() const { return m_num_local_faces; }
    [[nodiscard]] t8_locidx_t get_num_local_boundary_faces() const { return m_num_local_boundary_faces; }

    // ---- t8gpu_b200 extensions: the fused path -------------------------------------------------------------
    [[nodiscard]] t8b200_plan const* get_tile_plan() const { return m_plan; }

    /// One RK stage in one kernel (flux over all faces of this rank's elements + SSP_3RK_step{stage}): replaces
    /// kepes_compute_fluxes + reflective_boundary_condition + timestepping::SSP_3RK_step* of the reference's iterate().
    /// `in`: stage input, `prev`: U^n (ignored for stage 1), `out`: stage output.  speed_max_dev: device scalar that
    /// receives max(|uHat| + aHat) over the faces (for compute_timestep), or nullptr.  Asynchronous on `stream`; between
    /// stages the ranks must have finished the previous one (e.g. a 1-element ncclAllReduce on the same stream).
    void fused_stage(int stage, step_index_type in, step_index_type prev, step_index_type out, float_type delta_t,
                     float_type* speed_max_dev = nullptr, cudaStream_t stream = nullptr) {
      static_assert(nb_variables == T8B200_NVAR, "the fused kernels are the compressible-Euler ones (5 variables)");
      T8GPU_CUDA_CHECK_ERROR(t8b200_fused_stage(m_plan, stage, this->get_own_variables(in).data(),
                                                this->get_all_variables(in).data(), this->get_own_variables(prev).data(),
                                                this->get_own_variables(out).data(), this->get_own_volume(), delta_t,
                                                speed_max_dev, stream));
    }
    /// criteria[e] = sum over the interior faces of |rho_R - rho_L| / cbrt(volume) (estimate_gradient +
    /// compute_refinement_criteria of the example), without atomics and without touching the flux variables.
    void gradient_criteria(step_index_type step, variable_index_type rho, float_type* criteria_dev,
                           cudaStream_t stream = nullptr) {
      T8GPU_CUDA_CHECK_ERROR(t8b200_gradient_criteria(m_plan, this->get_own_variable(step, rho),
                                                      this->get_all_variables(step).data()[static_cast<int>(rho)],
                                                      this->get_own_volume(), criteria_dev, stream));
    }

   private:
    using Memory::resize;

    // precision dispatch onto the C ABI
    static cudaError_t t8b200_fused_stage(t8b200_plan const* p, int stage, float_type* const* in,
                                          float_type* const* const* in_all, float_type* const* prev,
                                          float_type* const* out, float_type const* vol, float_type dt,
                                          float_type* speed, cudaStream_t s) {
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_fused_stage_f64(p, stage, (double const* const*)in,
                                                               (double const* const* const*)in_all,
                                                               (double const* const*)prev, (double* const*)out,
                                                               (double const*)vol, dt, (double*)speed, s));
      else
        return static_cast<cudaError_t>(t8b200_fused_stage_f32(p, stage, (float const* const*)in,
                                                               (float const* const* const*)in_all,
                                                               (float const* const*)prev, (float* const*)out,
                                                               (float const*)vol, dt, (float*)speed, s));
    }
    static cudaError_t t8b200_gradient_criteria(t8b200_plan const* p, float_type const* rho, float_type* const* rho_all,
                                                float_type const* vol, float_type* out, cudaStream_t s) {
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_gradient_criteria_f64(p, (double const*)rho, (double const* const*)rho_all,
                                                                     (double const*)vol, (double*)out, s));
      else
        return static_cast<cudaError_t>(t8b200_gradient_criteria_f32(p, (float const*)rho, (float const* const*)rho_all,
                                                                     (float const*)vol, (float*)out, s));
    }
    static cudaError_t t8b200_partition_remap(int nvar, int64_t n, int cpe, int const* ranks, t8_locidx_t const* indices,
                                              float_type* const* un, float_type* const* const* uo, float_type* vn,
                                              float_type* const* vo, cudaStream_t s) {
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_partition_remap_f64(nvar, n, cpe, ranks, indices, (double* const*)un,
                                                                   (double const* const* const*)uo, (double*)vn,
                                                                   (double const* const*)vo, s));
      else
        return static_cast<cudaError_t>(t8b200_partition_remap_f32(nvar, n, cpe, ranks, indices, (float* const*)un,
                                                                   (float const* const* const*)uo, (float*)vn,
                                                                   (float const* const*)vo, s));
    }
    cudaError_t remap_adapt(t8_locidx_t const* device_map, t8_locidx_t n_new, step_index_type step, Memory& fresh) {
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_adapt_remap_f64(
            0, nb_variables, n_new, device_map, (double const* const*)this->get_own_variables(step).data(),
            (double* const*)fresh.get_own_variables(step).data(), (double const*)this->get_own_volume(),
            (double*)fresh.get_own_volume(), nullptr));
      else
        return static_cast<cudaError_t>(t8b200_adapt_remap_f32(
            0, nb_variables, n_new, device_map, (float const* const*)this->get_own_variables(step).data(),
            (float* const*)fresh.get_own_variables(step).data(), (float const*)this->get_own_volume(),
            (float*)fresh.get_own_volume(), nullptr));
    }

    /// fn(tree, scheme of the tree, element, running local index) for every local leaf, in SFC order
    template<typename Fn>
    static void for_each_leaf(t8_forest_t forest, Fn&& fn) {
      t8_locidx_t const num_trees = t8_forest_get_num_local_trees(forest);
      t8_locidx_t       idx       = 0;
      for (t8_locidx_t tree = 0; tree < num_trees; tree++) {
        t8_eclass_scheme_c* scheme = t8_forest_get_eclass_scheme(forest, t8_forest_get_tree_class(forest, tree));
        t8_locidx_t const   n      = t8_forest_get_tree_num_elements(forest, tree);
        for (t8_locidx_t i = 0; i < n; i++) fn(tree, scheme, t8_forest_get_element_in_tree(forest, tree, i), idx++);
      }
    }
    static std::vector<int> leaf_levels(t8_forest_t forest) {
      std::vector<int> levels(t8_forest_get_local_num_elements(forest));
      for_each_leaf(forest, [&](t8_locidx_t, t8_eclass_scheme_c* scheme, t8_element_t const* element, t8_locidx_t idx) {
        levels[idx] = scheme->t8_element_level(element);
      });
      return levels;
    }
    template<typename V>
    void ghost_exchange(V& per_element) {
      sc_array* wrapper = sc_array_new_data(per_element.data(), sizeof(typename V::value_type), per_element.size());
      t8_forest_ghost_exchange_data(m_forest, wrapper);
      sc_array_destroy(wrapper);
    }
    template<typename V>
    static void ship(t8_forest_t from, t8_forest_t to, V& data_from, V& data_to) {
      sc_array* in  = sc_array_new_data(data_from.data(), sizeof(typename V::value_type), data_from.size());
      sc_array* out = sc_array_new_data(data_to.data(), sizeof(typename V::value_type), data_to.size());
      t8_forest_partition_data(from, to, in, out);
      sc_array_destroy(in);
      sc_array_destroy(out);
    }
    void refresh_counts() {
      m_num_ghost_elements = t8_forest_get_num_ghosts(m_forest);
      m_num_local_elements = t8_forest_get_local_num_elements(m_forest);
    }
    /// device -> host, converted to double, written with the given stride
    void fetch(step_index_type step, variable_index_type variable, double* out, int stride) const {
      thrust::host_vector<float_type> h(m_num_local_elements);
      T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(h.data(), this->get_own_variable(step, variable),
                                        sizeof(float_type) * m_num_local_elements, cudaMemcpyDeviceToHost));
      for (t8_locidx_t i = 0; i < m_num_local_elements; i++) out[static_cast<size_t>(i) * stride] = static_cast<double>(h[i]);
    }

    struct UserData {
      thrust::host_vector<float_type>* element_refinement_criteria;
    };
    /// The reference's adapt rule (mesh_manager.inl:125-162): refine above b = 10 while below max_level; coarsen a
    /// family above min_level when the mean of its first FOUR criteria is below b (also in 3-D).
    static int adapt_callback_iteration(t8_forest_t, t8_forest_t forest_from, t8_locidx_t which_tree,
                                        t8_locidx_t lelement_id, t8_eclass_scheme_c* ts, int const is_family,
                                        int const, t8_element_t* elements[]) {
      auto* user = static_cast<UserData*>(t8_forest_get_user_data(forest_from));
      assert(user != nullptr);
      auto const&       crit   = *user->element_refinement_criteria;
      t8_locidx_t const level  = ts->t8_element_level(elements[0]);
      t8_locidx_t const first  = t8_forest_get_tree_element_offset(forest_from, which_tree) + lelement_id;
      float_type const  b      = static_cast<float_type>(10.0);
      if (level < max_level && crit[first] > b) return 1;
      if (level > min_level && is_family) {
        float_type mean = 0.0;
        for (int i = 0; i < 4; i++) mean += crit[first + i] / float_type{4.0};
        if (mean < b) return -1;
      }
      return 0;
    }

    sc_MPI_Comm      m_comm;
    int              m_rank{0};
    int              m_nb_ranks{1};
    t8_scheme_cxx_t* m_scheme;
    t8_cmesh_t       m_cmesh;
    t8_forest_t      m_forest;

    t8_locidx_t m_num_local_elements{0};
    t8_locidx_t m_num_ghost_elements{0};
    t8_locidx_t m_num_local_faces{0};
    t8_locidx_t m_num_local_boundary_faces{0};
    t8_locidx_t m_num_x_faces{0};

    std::vector<int>                   m_ranks;
    std::vector<t8_locidx_t>           m_indices;
    thrust::device_vector<int>         m_device_ranks;
    thrust::device_vector<t8_locidx_t> m_device_indices;
    thrust::device_vector<t8_locidx_t> m_device_face_neighbors;  ///< interior pairs, then boundary singles
    thrust::device_vector<float_type>  m_device_face_normals;    ///< dim components per face, interior then boundary
    thrust::device_vector<float_type>  m_device_face_area;

    thrust::host_vector<float_type> m_element_refinement_criteria;
    UserData                        m_user_data{};
    t8b200_plan*                    m_plan{nullptr};
  };

}  // namespace t8gpu

#endif  // T8GPU_B200_MESH_MESH_MANAGER_H
