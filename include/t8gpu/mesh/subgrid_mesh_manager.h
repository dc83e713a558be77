/// @file subgrid_mesh_manager.h
/// @brief SubgridMeshManager and SubgridMeshConnectivityAccessor: source-compatible with
///        t8gpu/mesh/subgrid_mesh_manager.h:29-509 (same template parameters, member names, accessor getters).
///
/// Same as the reference (needed for bit-identical connectivity): the t8code calls that own the forest, the adapt rule
/// (threshold 0.02, levels 1..6), the face enumeration order and the canonical form of a face record -- left element
/// never coarser than the right one, `level_difference` 0 or -1, `neighbor_offset` = anchor cell of the face inside the
/// right element (t8gpu/mesh/subgrid_mesh_manager.inl:560-961).
///
/// Different behind the same interface: one leaf walk for everything, partition-boundary faces of the lower rank are
/// recorded too ("x" faces) and the cell-level tile plan of the fused stage kernel is built
/// (t8b200_subgrid_plan_create); adapt() / partition() remap on the device straight into the new allocation
/// (t8b200_adapt_remap / t8b200_partition_remap); fused_stage() / refinement_criteria() expose the fused kernels.
#ifndef T8GPU_B200_MESH_SUBGRID_MESH_MANAGER_H
#define T8GPU_B200_MESH_SUBGRID_MESH_MANAGER_H

#include <t8.h>
#include <t8_cmesh.h>
#include <t8_forest/t8_forest.h>
#include <t8_forest/t8_forest_io.h>
#include <t8_forest/t8_forest_partition.h>
#include <t8gpu/memory/subgrid_memory_manager.h>
#include <t8gpu/utils/cuda.h>
#include <t8gpu/utils/meta.h>
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

  /// Kernel-argument POD with the face connectivity of one rank (subgrid_mesh_manager.h:29-216).
  template<typename float_type, typename SubgridType>
  class SubgridMeshConnectivityAccessor {
    template<typename VT, typename ST, typename SubgridType_>
    friend class SubgridMeshManager;
    constexpr static int dim = SubgridType::rank;

   public:
    SubgridMeshConnectivityAccessor(SubgridMeshConnectivityAccessor const&)            = default;
    SubgridMeshConnectivityAccessor& operator=(SubgridMeshConnectivityAccessor const&) = default;

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
      for (int k = 0; k < dim; k++) n[k] = m_face_normals[dim * face_idx + k];
      return n;
    }
    [[nodiscard]] __device__ inline std::array<float_type, dim> get_boundary_face_normal(int face_idx) const {
      return get_face_normal(m_num_local_faces + face_idx);
    }
    /// 0, or -1 when the right element is one level coarser than the left one.
    [[nodiscard]] __device__ inline t8_locidx_t get_face_level_difference(int face_idx) {
      return m_face_level_difference[face_idx];
    }
    /// anchor cell of the face inside the right element.
    [[nodiscard]] __device__ inline std::array<t8_locidx_t, SubgridType::rank> get_face_neighbor_offset(int f_idx) {
      std::array<t8_locidx_t, SubgridType::rank> o{};
      for (int k = 0; k < dim; k++) o[k] = m_face_neighbor_offset[dim * f_idx + k];
      return o;
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
    [[nodiscard]] __host__ t8_locidx_t const* face_level_difference() const { return m_face_level_difference; }
    [[nodiscard]] __host__ t8_locidx_t const* face_neighbor_offset() const { return m_face_neighbor_offset; }

   private:
    int const*         m_ranks;
    t8_locidx_t const* m_indices;
    t8_locidx_t const* m_face_neighbors;
    t8_locidx_t const* m_face_level_difference;
    t8_locidx_t const* m_face_neighbor_offset;
    float_type const*  m_face_normals;
    float_type const*  m_face_surfaces;
    t8_locidx_t        m_num_local_faces;
    t8_locidx_t        m_num_local_boundary_faces;

    SubgridMeshConnectivityAccessor(int const* ranks, t8_locidx_t const* indices, t8_locidx_t const* face_neighbors,
                                    t8_locidx_t const* level_difference, t8_locidx_t const* neighbor_offset,
                                    float_type const* face_normals, float_type const* face_surfaces, t8_locidx_t nf,
                                    t8_locidx_t nb)
        : m_ranks{ranks}, m_indices{indices}, m_face_neighbors{face_neighbors},
          m_face_level_difference{level_difference}, m_face_neighbor_offset{neighbor_offset},
          m_face_normals{face_normals}, m_face_surfaces{face_surfaces}, m_num_local_faces{nf},
          m_num_local_boundary_faces{nb} {}
  };

  namespace detail {
    /// every cell of element e <- per-element value (initialisation from one value per element)
    template<typename float_type, int nvar, int S>
    __global__ void broadcast_to_cells(float_type const* per_element, std::array<float_type*, nvar> cells, int64_t n) {
      int64_t const g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
      if (g >= n * S) return;
      for (int k = 0; k < nvar; k++) cells[k][g] = per_element[k * n + g / S];
    }
  }  // namespace detail

  /// Owns a t8code forest whose elements each carry a Cartesian subgrid of cells (subgrid_mesh_manager.h:265-509).
  template<typename VariableType, typename StepType, typename SubgridType>
  class SubgridMeshManager : public SubgridMemoryManager<VariableType, StepType, SubgridType> {
    using Memory = SubgridMemoryManager<VariableType, StepType, SubgridType>;

   public:
    using float_type                  = typename variable_traits<VariableType>::float_type;
    using variable_index_type         = typename variable_traits<VariableType>::index_type;
    static constexpr int nb_variables = variable_traits<VariableType>::nb_variables;
    static constexpr int dim          = SubgridType::rank;
    using step_index_type             = typename step_traits<StepType>::index_type;
    static constexpr size_t nb_steps  = step_traits<StepType>::nb_steps;

    static constexpr t8_locidx_t min_level = 1;
    static constexpr t8_locidx_t max_level = 6;

    /// Takes ownership of cmesh and forest; collective.  Volumes are set here (subgrid_mesh_manager.inl:66-84).
    SubgridMeshManager(sc_MPI_Comm comm, t8_scheme_cxx_t* scheme, t8_cmesh_t cmesh, t8_forest_t forest)
        : Memory{static_cast<size_t>(t8_forest_get_local_num_elements(forest)), comm},
          m_comm{comm}, m_scheme{scheme}, m_cmesh{cmesh}, m_forest{forest} {
      static_assert(dim == 2 || dim == 3, "quad or hex forests");
      static_assert(SubgridType::size == (dim == 3 ? 64 : 16), "the fused kernels are built for Subgrid<4,4[,4]>");
      MPI_Comm_size(m_comm, &m_nb_ranks);
      MPI_Comm_rank(m_comm, &m_rank);
      refresh_counts();
      m_user_data.element_refinement_criteria = &m_element_refinement_criteria;
      t8_forest_set_user_data(m_forest, &m_user_data);
      m_element_refinement_criteria.resize(m_num_local_elements);
      m_host_volume = leaf_volumes(m_forest);
      this->set_volume(m_host_volume);
      compute_connectivity_information();
    }
    ~SubgridMeshManager() {
      t8b200_subgrid_plan_destroy(m_plan);
      t8_forest_unref(&m_forest);
      t8_cmesh_destroy(&m_cmesh);
    }
    SubgridMeshManager(SubgridMeshManager const&)            = delete;
    SubgridMeshManager& operator=(SubgridMeshManager const&) = delete;

    /// func(MemoryAccessorOwn& host_variables, t8_forest_t, tree, element, index) gives ONE value per element and
    /// variable; every cell of the element receives it (subgrid_mesh_manager.h:300-312).
    template<typename Func>
    void initialize_variables(Func func) {
      size_t const n = static_cast<size_t>(m_num_local_elements);
      thrust::host_vector<float_type>       host(n * nb_variables);
      std::array<float_type*, nb_variables> rows{};
      for (int k = 0; k < nb_variables; k++) rows[k] = host.data() + k * n;
      MemoryAccessorOwn<VariableType> host_accessor{rows};
      for_each_leaf(m_forest, [&](t8_locidx_t tree, t8_eclass_scheme_c*, t8_element_t const* element, t8_locidx_t idx) {
        func(host_accessor, m_forest, tree, element, idx);
      });
      if (n == 0) return;
      thrust::device_vector<float_type>     per_element = host;
      std::array<float_type*, nb_variables> cells{};
      auto                                  own = this->get_own_variables(static_cast<step_index_type>(0));
      for (int k = 0; k < nb_variables; k++) cells[k] = own.data()[k];
      int64_t const total = static_cast<int64_t>(n) * SubgridType::size;
      detail::broadcast_to_cells<float_type, nb_variables, SubgridType::size>
          <<<static_cast<unsigned>((total + 255) / 256), 256>>>(thrust::raw_pointer_cast(per_element.data()), cells,
                                                                static_cast<int64_t>(n));
      T8GPU_CUDA_CHECK_LAST_ERROR();
      T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize());
    }

    /// t8code adapt (+ balance, ghosts), then device remap of the variables of `step` and of the volumes.  Collective.
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

      std::vector<int> const old_levels = leaf_levels(m_forest), new_levels = leaf_levels(adapted);
      constexpr int          nb_children = 1 << dim;
      t8_locidx_t const      n_old = static_cast<t8_locidx_t>(old_levels.size()),
                        n_new      = static_cast<t8_locidx_t>(new_levels.size());
      thrust::host_vector<t8_locidx_t> map(n_new + 1);
      t8_locidx_t                      o = 0, n = 0;
      while (o < n_old && n < n_new) {
        if (old_levels[o] < new_levels[n]) {
          for (int c = 0; c < nb_children; c++) map[n + c] = o;
          o += 1;
          n += nb_children;
        } else if (old_levels[o] > new_levels[n]) {
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
      fetch_volumes();
    }

    /// t8code repartition; every rank pulls its new elements (all cells + volume) from the ranks that held them.
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
      T8GPU_CUDA_CHECK_ERROR(remap_partition(thrust::raw_pointer_cast(d_ranks.data()),
                                             thrust::raw_pointer_cast(d_indices.data()), n_new, step, fresh));
      T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize());
      MPI_Barrier(m_comm);
      static_cast<Memory&>(*this) = std::move(fresh);
      m_element_refinement_criteria.resize(n_new);

      t8_forest_set_user_data(partitioned, &m_user_data);
      t8_forest_unref(&m_forest);
      m_forest = partitioned;
      refresh_counts();
      fetch_volumes();
    }

    /// Rebuilds owner tables, face arrays (bit for bit the reference's) and the cell-level tile plan.  Collective.
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

      struct Faces {
        std::vector<t8_locidx_t> nbr, level_diff, offset;
        std::vector<float_type>  normals, areas;
      } in, x;
      std::vector<t8_locidx_t> bnbr;
      std::vector<float_type>  bnrm, barea;
      constexpr int            E = SubgridType::template extent<0>;

      // one face record in canonical form: left = the finer (or equal) element
      auto add_face = [&](Faces& out, int face, int share, t8_locidx_t tree, t8_element_t const* element,
                          t8_eclass_scheme_c* scheme, t8_locidx_t idx, t8_element_t const* neighbor,
                          t8_eclass_scheme_c* nscheme, t8_locidx_t nidx) {
        int const level = scheme->t8_element_level(element), nlevel = nscheme->t8_element_level(neighbor);
        int const axis = face / 2, plus = face % 2;
        double    normal[3] = {0.0, 0.0, 0.0};
        t8_forest_element_face_normal(m_forest, tree, element, face, normal);
        std::array<int, dim> offset{};
        bool const           swap = nlevel > level;   // neighbour finer: it becomes the left element
        if (nlevel == level) {
          offset[axis] = plus ? 0 : E - 1;
        } else {
          // the fine side's child id places the face inside the coarse element
          int const child = swap ? nscheme->t8_element_child_id(neighbor) : scheme->t8_element_child_id(element);
          for (int d = 0; d < dim; d++) offset[d] = E / 2 * ((child >> d) & 1);
          offset[axis] = (plus != swap) ? 0 : E - 1;
        }
        out.level_diff.push_back(swap ? level - nlevel : nlevel - level);
        for (int d = 0; d < dim; d++) out.offset.push_back(offset[d]);
        out.nbr.push_back(swap ? nidx : idx);
        out.nbr.push_back(swap ? idx : nidx);
        for (int d = 0; d < dim; d++) out.normals.push_back(static_cast<float_type>(swap ? -normal[d] : normal[d]));
        out.areas.push_back(static_cast<float_type>(t8_forest_element_face_area(m_forest, tree, element, face)) /
                            static_cast<float_type>(share));
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
          for (int i = 0; i < num_neighbors; i++) {
            if (ids[i] < m_num_local_elements) continue;
            add_face(m_rank < m_ranks[ids[i]] ? in : x, face, num_neighbors, tree, element, scheme, idx, neighbors[i],
                     nscheme, ids[i]);
          }
          if (num_neighbors == 1 && ids[0] < m_num_local_elements) {
            bool const lower = ids[0] > idx;
            bool const finer = ids[0] < idx && nscheme->t8_element_level(neighbors[0]) < scheme->t8_element_level(element);
            if (lower || finer) add_face(in, face, 1, tree, element, scheme, idx, neighbors[0], nscheme, ids[0]);
          }
          if (num_neighbors == 0) {
            bnbr.push_back(idx);
            double normal[3] = {0.0, 0.0, 0.0};
            t8_forest_element_face_normal(m_forest, tree, element, face, normal);
            for (int d = 0; d < dim; d++) bnrm.push_back(static_cast<float_type>(normal[d]));
            barea.push_back(static_cast<float_type>(t8_forest_element_face_area(m_forest, tree, element, face)));
          }
          if (neighbors) {
            nscheme->t8_element_destroy(num_neighbors, neighbors);
            T8_FREE(neighbors);
          }
          T8_FREE(dual_faces);
          T8_FREE(ids);
        }
      });
      m_num_local_faces          = static_cast<t8_locidx_t>(in.areas.size());
      m_num_local_boundary_faces = static_cast<t8_locidx_t>(barea.size());
      t8_locidx_t const n_x      = static_cast<t8_locidx_t>(x.areas.size());
      in.nbr.insert(in.nbr.end(), bnbr.begin(), bnbr.end());
      in.normals.insert(in.normals.end(), bnrm.begin(), bnrm.end());
      in.areas.insert(in.areas.end(), barea.begin(), barea.end());
      m_device_face_neighbors.assign(in.nbr.begin(), in.nbr.end());
      m_device_face_normals.assign(in.normals.begin(), in.normals.end());
      m_device_face_area.assign(in.areas.begin(), in.areas.end());
      m_device_face_level_difference.assign(in.level_diff.begin(), in.level_diff.end());
      m_device_face_neighbor_offset.assign(in.offset.begin(), in.offset.end());

      if (static_cast<t8_locidx_t>(m_host_volume.size()) != m_num_local_elements) fetch_volumes();
      t8b200_subgrid_plan_destroy(m_plan);
      m_plan = nullptr;
      T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_plan_create(
          &m_plan, sizeof(float_type) == 8, dim, m_num_local_elements, m_num_ghost_elements, m_num_local_faces,
          m_num_local_boundary_faces, in.nbr.data(), in.normals.data(), in.areas.data(), in.level_diff.data(),
          in.offset.data(), m_host_volume.data(), m_ranks.data(), m_indices.data(), n_x, x.nbr.data(),
          x.normals.data(), x.areas.data(), x.level_diff.data(), x.offset.data()));
    }

    // ---- output (subgrid_mesh_manager.h:355-446, .inl:1007-1208)

    /// Cell data of one variable, one value per CELL, written on the forest refined log2(extent) more times so that
    /// every cell is a leaf (the reference's scheme, subgrid_mesh_manager.inl:1066-1124).  The permutation to that
    /// forest's leaf order and the widening to double run on the device (t8b200_subgrid_z_order), one copy to the host.
    void save_variable_to_vtk(step_index_type step, variable_index_type variable, std::string const& prefix) const {
      size_t const                  cells = static_cast<size_t>(m_num_local_elements) * SubgridType::size;
      thrust::device_vector<double> z_ordered(cells);
      float_type const* const       from = static_cast<float_type const*>(this->get_own_variable(step, variable));
      if constexpr (sizeof(float_type) == 8)
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_z_order_f64(dim, m_num_local_elements, (double const*)from,
                                                          thrust::raw_pointer_cast(z_ordered.data()), nullptr));
      else
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_z_order_f32(dim, m_num_local_elements, (float const*)from,
                                                          thrust::raw_pointer_cast(z_ordered.data()), nullptr));
      thrust::host_vector<double> host = z_ordered;

      constexpr int levels = meta::log2_v<SubgridType::template extent<0>>;
      t8_forest_t   fine   = m_forest;
      for (int l = 0; l < levels; l++) {   // every element -> its 2^dim children, `levels` times
        t8_forest_t next{};
        t8_forest_init(&next);
        if (l == 0) t8_forest_ref(fine);   // set_adapt consumes one reference of its source; m_forest stays ours
        t8_forest_set_adapt(next, fine, refine_everything, false);
        t8_forest_commit(next);
        fine = next;
      }
      t8_vtk_data_field_t field{};
      field.type = T8_VTK_SCALAR;
      std::strncpy(field.description, "variables", BUFSIZ - 1);
      field.data = host.data();
      t8_forest_write_vtk_ext(fine, prefix.c_str(), true, true, true, true, false, false, false, 1, &field);
      if (levels > 0) t8_forest_unref(&fine);
    }

    /// The element mesh alone (subgrid_mesh_manager.inl:1126-1142).
    void save_mesh_to_vtk(std::string const& prefix) const {
      t8_forest_write_vtk_ext(m_forest, prefix.c_str(), true, true, true, true, false, false, false, 0, nullptr);
    }

    /// Host copy of a variable, owning wrapper around t8_vtk_data_field_t (subgrid_mesh_manager.h:388-417).
    class HostVariableInfo {
      friend SubgridMeshManager;
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
      t8_vtk_data_field_t       m_field{};
    };

    /// As the reference (subgrid_mesh_manager.inl:1144-1161): the first num_local_elements values of the variable's
    /// array, widened to double.
    [[nodiscard]] HostVariableInfo get_host_scalar_variable(step_index_type step, variable_index_type variable,
                                                            std::string const& name) const {
      auto data = std::make_unique<double[]>(static_cast<size_t>(m_num_local_elements));
      fetch_head(step, variable, data.get(), 1);
      return HostVariableInfo{T8_VTK_SCALAR, std::move(data), name};
    }
    /// Three variables interleaved per entry (subgrid_mesh_manager.inl:1163-1183).
    [[nodiscard]] HostVariableInfo get_host_vector_variable(step_index_type step, std::array<variable_index_type, 3> variable,
                                                            std::string const& name) const {
      auto data = std::make_unique<double[]>(3 * static_cast<size_t>(m_num_local_elements));
      for (int c = 0; c < 3; c++) fetch_head(step, variable[c], data.get() + c, 3);
      return HostVariableInfo{T8_VTK_VECTOR, std::move(data), name};
    }
    /// The reference's body is commented out (subgrid_mesh_manager.inl:1185-1208): the fields are per element while the
    /// data is per cell.  Kept callable with the same effect: nothing is written.
    void save_variables_to_vtk(std::vector<HostVariableInfo> host_variables, std::string const& prefix) const {
      (void)host_variables;
      (void)prefix;
    }

    [[nodiscard]] SubgridMeshConnectivityAccessor<float_type, SubgridType> get_connectivity_information() const {
      return {thrust::raw_pointer_cast(m_device_ranks.data()),
              thrust::raw_pointer_cast(m_device_indices.data()),
              thrust::raw_pointer_cast(m_device_face_neighbors.data()),
              thrust::raw_pointer_cast(m_device_face_level_difference.data()),
              thrust::raw_pointer_cast(m_device_face_neighbor_offset.data()),
              thrust::raw_pointer_cast(m_device_face_normals.data()),
              thrust::raw_pointer_cast(m_device_face_area.data()),
              m_num_local_faces,
              m_num_local_boundary_faces};
    }
    [[nodiscard]] t8_locidx_t get_num_local_elements() const { return m_num_local_elements; }
    [[nodiscard]] t8_locidx_t get_num_ghost_elements() const { return m_num_ghost_elements; }
    [[nodiscard]] t8_locidx_t get_num_local_faces() const { return m_num_local_faces; }
    [[nodiscard]] t8_locidx_t get_num_local_boundary_faces() const { return m_num_local_boundary_faces; }

    // ---- t8gpu_b200 extensions: the fused path -------------------------------------------------------------
    [[nodiscard]] t8b200_subgrid_plan const* get_tile_plan() const { return m_plan; }

    /// One RK stage in one kernel: replaces compute_inner_fluxes + compute_boundary_fluxes + compute_outer_fluxes +
    /// timestepping::subgrid::SSP_3RK_step{stage} of the reference's iterate() (examples/subgrid/solver.inl:156-194).
    void fused_stage(int stage, step_index_type in, step_index_type prev, step_index_type out, float_type delta_t,
                     cudaStream_t stream = nullptr) {
      static_assert(nb_variables == T8B200_NVAR, "the fused kernels are the compressible-Euler ones (5 variables)");
      auto i = this->get_own_variables(in);
      auto a = this->get_all_variables(in);
      auto p = this->get_own_variables(prev);
      auto o = this->get_own_variables(out);
      if constexpr (sizeof(float_type) == 8)
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_fused_stage_f64(
            m_plan, stage, (double const* const*)i.data(), (double const* const* const*)a.data(),
            (double const* const*)p.data(), (double* const*)o.data(), (double const*)this->get_own_volume(), delta_t,
            stream));
      else
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_fused_stage_f32(
            m_plan, stage, (float const* const*)i.data(), (float const* const* const*)a.data(),
            (float const* const*)p.data(), (float* const*)o.data(), (float const*)this->get_own_volume(), delta_t,
            stream));
    }
    /// H1 seminorm of one variable per element / volume (compute_refinement_criteria of the subgrid example).
    void refinement_criteria(step_index_type step, variable_index_type variable, float_type* criteria_dev,
                             cudaStream_t stream = nullptr) {
      auto own = this->get_own_variables(step);
      if constexpr (sizeof(float_type) == 8)
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_criteria_f64(dim, m_num_local_elements,
                                                           (double const*)own.data()[static_cast<int>(variable)],
                                                           (double const*)this->get_own_volume(), (double*)criteria_dev,
                                                           stream));
      else
        T8GPU_CUDA_CHECK_ERROR(t8b200_subgrid_criteria_f32(dim, m_num_local_elements,
                                                           (float const*)own.data()[static_cast<int>(variable)],
                                                           (float const*)this->get_own_volume(), (float*)criteria_dev,
                                                           stream));
    }

   private:
    using Memory::resize;

    cudaError_t remap_adapt(t8_locidx_t const* device_map, t8_locidx_t n_new, step_index_type step, Memory& fresh) {
      auto o = this->get_own_variables(step);
      auto n = fresh.get_own_variables(step);
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_adapt_remap_f64(dim, nb_variables, n_new, device_map,
                                                               (double const* const*)o.data(), (double* const*)n.data(),
                                                               (double const*)this->get_own_volume(),
                                                               (double*)fresh.get_own_volume(), nullptr));
      else
        return static_cast<cudaError_t>(t8b200_adapt_remap_f32(dim, nb_variables, n_new, device_map,
                                                               (float const* const*)o.data(), (float* const*)n.data(),
                                                               (float const*)this->get_own_volume(),
                                                               (float*)fresh.get_own_volume(), nullptr));
    }
    cudaError_t remap_partition(int const* ranks, t8_locidx_t const* indices, t8_locidx_t n_new, step_index_type step,
                                Memory& fresh) {
      auto o = this->get_all_variables(step);
      auto n = fresh.get_own_variables(step);
      if constexpr (sizeof(float_type) == 8)
        return static_cast<cudaError_t>(t8b200_partition_remap_f64(
            nb_variables, n_new, SubgridType::size, ranks, indices, (double* const*)n.data(),
            (double const* const* const*)o.data(), (double*)fresh.get_own_volume(),
            (double const* const*)this->get_all_volume(), nullptr));
      else
        return static_cast<cudaError_t>(t8b200_partition_remap_f32(
            nb_variables, n_new, SubgridType::size, ranks, indices, (float* const*)n.data(),
            (float const* const* const*)o.data(), (float*)fresh.get_own_volume(),
            (float const* const*)this->get_all_volume(), nullptr));
    }

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
    static thrust::host_vector<float_type> leaf_volumes(t8_forest_t forest) {
      thrust::host_vector<float_type> v(t8_forest_get_local_num_elements(forest));
      for_each_leaf(forest, [&](t8_locidx_t tree, t8_eclass_scheme_c*, t8_element_t const* element, t8_locidx_t idx) {
        v[idx] = static_cast<float_type>(t8_forest_element_volume(forest, tree, element));
      });
      return v;
    }
    /// t8code adapt callback that refines every element (the cells of a subgrid as leaves, for output)
    static int refine_everything(t8_forest_t, t8_forest_t, t8_locidx_t, t8_locidx_t, t8_eclass_scheme_c*, int const,
                                 int const, t8_element_t*[]) {
      return 1;
    }
    /// first num_local_elements values of a variable's array -> out[i * stride], widened to double
    void fetch_head(step_index_type step, variable_index_type variable, double* out, int stride) const {
      thrust::host_vector<float_type> h(m_num_local_elements);
      if (m_num_local_elements > 0)
        T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(h.data(), static_cast<float_type const*>(this->get_own_variable(step, variable)),
                                          sizeof(float_type) * m_num_local_elements, cudaMemcpyDeviceToHost));
      for (t8_locidx_t i = 0; i < m_num_local_elements; i++) out[static_cast<size_t>(i) * stride] = static_cast<double>(h[i]);
    }
    /// host copy of the (remapped) volumes: the plan needs them for the areas of the faces between cells
    void fetch_volumes() {
      m_host_volume.resize(m_num_local_elements);
      if (m_num_local_elements > 0)
        T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_host_volume.data(), this->get_own_volume(),
                                          sizeof(float_type) * m_num_local_elements, cudaMemcpyDeviceToHost));
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

    struct UserData {
      thrust::host_vector<float_type>* element_refinement_criteria;
    };
    /// The reference's adapt rule (subgrid_mesh_manager.inl:196-234): refine above b = 0.02 while below max_level;
    /// coarsen a family above min_level when the mean of its first FOUR criteria is below b.
    static int adapt_callback_iteration(t8_forest_t, t8_forest_t forest_from, t8_locidx_t which_tree,
                                        t8_locidx_t lelement_id, t8_eclass_scheme_c* ts, int const is_family,
                                        int const, t8_element_t* elements[]) {
      auto* user = static_cast<UserData*>(t8_forest_get_user_data(forest_from));
      assert(user != nullptr);
      auto const&       crit  = *user->element_refinement_criteria;
      t8_locidx_t const level = ts->t8_element_level(elements[0]);
      t8_locidx_t const first = t8_forest_get_tree_element_offset(forest_from, which_tree) + lelement_id;
      float_type const  b     = static_cast<float_type>(0.02);
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

    std::vector<int>                   m_ranks;
    std::vector<t8_locidx_t>           m_indices;
    thrust::device_vector<int>         m_device_ranks;
    thrust::device_vector<t8_locidx_t> m_device_indices;
    thrust::device_vector<t8_locidx_t> m_device_face_neighbors;
    thrust::device_vector<t8_locidx_t> m_device_face_level_difference;
    thrust::device_vector<t8_locidx_t> m_device_face_neighbor_offset;
    thrust::device_vector<float_type>  m_device_face_normals;
    thrust::device_vector<float_type>  m_device_face_area;

    thrust::host_vector<float_type> m_host_volume;
    thrust::host_vector<float_type> m_element_refinement_criteria;
    UserData                        m_user_data{};
    t8b200_subgrid_plan*            m_plan{nullptr};
  };

}  // namespace t8gpu

#endif  // T8GPU_B200_MESH_SUBGRID_MESH_MANAGER_H
