/// @file subgrid_memory_manager.h
/// @brief Subgrid<extents...>, its Accessor, SubgridMemoryAccessor{Own,All} and SubgridMemoryManager;
///        source-compatible with t8gpu/memory/subgrid_memory_manager.h:35-555.
#ifndef T8GPU_B200_MEMORY_SUBGRID_MEMORY_MANAGER_H
#define T8GPU_B200_MEMORY_SUBGRID_MEMORY_MANAGER_H

#include <t8gpu/memory/memory_manager.h>
#include <t8gpu/memory/shared_device_vector.h>
#include <t8gpu/utils/meta.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <array>
#include <tuple>
#include <type_traits>
#include <utility>

namespace t8gpu {

  /// A structured block of cells inside one mesh element, column major (first index fastest):
  /// Subgrid<4,4,4>::flat_index(i,j,k) = i + 4 j + 16 k.
  template<int... extents>
  struct Subgrid {
    static constexpr int rank = sizeof...(extents);
    static constexpr int size = (extents * ...);
    template<int dim>
    static constexpr int extent = meta::argpack_at_v<dim, extents...>;
    template<int i>
    static constexpr int stride = meta::argpack_mul_to_v<i, extents...>;

    template<typename... Ts>
    __host__ __device__ static constexpr inline int flat_index(Ts... is) {
      static_assert(sizeof...(Ts) == rank, "one index per subgrid dimension");
      return flat_index_impl(std::index_sequence_for<Ts...>{}, is...);
    }

    static constexpr dim3 block_size = {extents...};

    /// View of one variable: accessor(e, i, j, k) -> cell (i,j,k) of element e.
    template<typename float_type>
    class Accessor {
     public:
      Accessor(Accessor const&)            = default;
      Accessor& operator=(Accessor const&) = default;

      template<typename... Ts>
      [[nodiscard]] inline __device__
          std::enable_if_t<(sizeof...(Ts) == rank) && std::conjunction_v<std::is_integral<Ts>...>, float_type&>
          operator()(size_t e_idx, Ts... is) {
        return m_data[e_idx * size + flat_index(is...)];
      }
      template<typename... Ts>
      [[nodiscard]] inline __device__
          std::enable_if_t<(sizeof...(Ts) == rank) && std::conjunction_v<std::is_integral<Ts>...>, float_type const&>
          operator()(size_t e_idx, Ts... is) const {
        return m_data[e_idx * size + flat_index(is...)];
      }
      __host__ __device__ explicit operator float_type*() { return m_data; }
      __host__ __device__ explicit operator float_type const*() const { return m_data; }

     private:
      __device__ __host__ Accessor(float_type const* data) : m_data{const_cast<float_type*>(data)} {}
      float_type* m_data;

      template<typename VariableType, typename SubgridType>
      friend class SubgridMemoryAccessorOwn;
      template<typename VariableType, typename SubgridType>
      friend class SubgridMemoryAccessorAll;
      template<typename VariableType, typename StepType, typename SubgridType>
      friend class SubgridMemoryManager;
    };
    template<typename float_type>
    using accessor_type = Accessor<float_type>;

   private:
    template<size_t... I, typename... Ts>
    __host__ __device__ static constexpr inline int flat_index_impl(std::index_sequence<I...>, Ts... is) {
      return ((stride<I> * static_cast<int>(is)) + ...);
    }
  };

  /// Variables of the elements owned by this rank: `acc.get(Rho)(e, i, j, k)` (subgrid_memory_manager.h:178-276).
  template<typename VariableType, typename SubgridType>
  class SubgridMemoryAccessorOwn {
    template<typename VT, typename ST, typename SubgridType_>
    friend class SubgridMemoryManager;
    template<typename VT, typename ST, typename SubgridType_>
    friend class SubgridMeshManager;

   public:
    using variable_index_type            = typename variable_traits<VariableType>::index_type;
    using float_type                     = typename variable_traits<VariableType>::float_type;
    constexpr static size_t nb_variables = variable_traits<VariableType>::nb_variables;
    using view_type                      = typename SubgridType::template accessor_type<float_type>;

    SubgridMemoryAccessorOwn(SubgridMemoryAccessorOwn const&)            = default;
    SubgridMemoryAccessorOwn& operator=(SubgridMemoryAccessorOwn const&) = default;

    template<typename T>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        meta::is_explicitly_convertible_to_v<T, variable_index_type>, view_type>
    get(T i) {
      return view_type{m_pointers[static_cast<variable_index_type>(i)]};
    }
    template<typename T>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        meta::is_explicitly_convertible_to_v<T, variable_index_type>, view_type const>
    get(T i) const {
      return view_type{m_pointers[static_cast<variable_index_type>(i)]};
    }
    template<typename... Ts>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        (sizeof...(Ts) > 1) &&
            meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
            meta::all_same_v<Ts...>,
        std::array<view_type, sizeof...(Ts)>>
    get(Ts... is) {
      return {get(static_cast<variable_index_type>(is))...};
    }
    template<typename... Ts>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        (sizeof...(Ts) > 1) &&
            meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
            meta::all_same_v<Ts...>,
        std::array<view_type const, sizeof...(Ts)>>
    get(Ts... is) const {
      return {get(static_cast<variable_index_type>(is))...};
    }

   private:
    std::array<float_type*, nb_variables> m_pointers;

   public:
    /// host-side view of the pointer array, as the C ABI (t8gpu_b200.h) takes it
    [[nodiscard]] __host__ float_type* const* data() const { return m_pointers.data(); }

   private:
    // (constrained so that it never competes with the copy constructor for a non-const accessor)
    template<typename Container,
             typename = std::enable_if_t<!std::is_same_v<std::decay_t<Container>, SubgridMemoryAccessorOwn>>>
    SubgridMemoryAccessorOwn(Container&& array) : m_pointers(std::forward<Container>(array)) {}
  };

  /// Variables of every rank: `acc.get(rank, Rho)(e, i, j, k)` (subgrid_memory_manager.h:310-411).
  template<typename VariableType, typename SubgridType>
  class SubgridMemoryAccessorAll {
    template<typename VT, typename ST, typename SubgridType_>
    friend class SubgridMemoryManager;
    template<typename VT, typename ST, typename SubgridType_>
    friend class SubgridMeshManager;

   public:
    using variable_index_type            = typename variable_traits<VariableType>::index_type;
    using float_type                     = typename variable_traits<VariableType>::float_type;
    constexpr static size_t nb_variables = variable_traits<VariableType>::nb_variables;
    using view_type                      = typename SubgridType::template accessor_type<float_type>;

    SubgridMemoryAccessorAll(SubgridMemoryAccessorAll const&)            = default;
    SubgridMemoryAccessorAll& operator=(SubgridMemoryAccessorAll const&) = default;

    template<typename T>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        meta::is_explicitly_convertible_to_v<T, variable_index_type>, view_type>
    get(int rank, T i) {
      return view_type{m_pointers[static_cast<variable_index_type>(i)][rank]};
    }
    template<typename T>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        meta::is_explicitly_convertible_to_v<T, variable_index_type>, view_type const>
    get(int rank, T i) const {
      return view_type{m_pointers[static_cast<variable_index_type>(i)][rank]};
    }
    template<typename... Ts>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        (sizeof...(Ts) > 1) &&
            meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
            meta::all_same_v<Ts...>,
        std::array<view_type, sizeof...(Ts)>>
    get(int rank, Ts... is) {
      return {get(rank, static_cast<variable_index_type>(is))...};
    }
    template<typename... Ts>
    [[nodiscard]] __device__ __host__ inline std::enable_if_t<
        (sizeof...(Ts) > 1) &&
            meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
            meta::all_same_v<Ts...>,
        std::array<view_type const, sizeof...(Ts)>>
    get(int rank, Ts... is) const {
      return {get(rank, static_cast<variable_index_type>(is))...};
    }

   private:
    std::array<float_type* const*, nb_variables> m_pointers;

   public:
    /// host-side view of the table array, as the C ABI (t8gpu_b200.h) takes it
    [[nodiscard]] __host__ float_type* const* const* data() const { return m_pointers.data(); }

   private:
    template<typename Container,
             typename = std::enable_if_t<!std::is_same_v<std::decay_t<Container>, SubgridMemoryAccessorAll>>>
    SubgridMemoryAccessorAll(Container&& array) : m_pointers(std::forward<Container>(array)) {}
  };

  /// nb_variables x nb_steps arrays of nb_elements * Subgrid::size cells + a per-element volume vector
  /// (subgrid_memory_manager.h:424-555).  As in the reference, resize() does not touch the volume; set_volume does.
  template<typename VariableType, typename StepType, typename SubgridType>
  class SubgridMemoryManager {
   public:
    using float_type                     = typename variable_traits<VariableType>::float_type;
    using variable_index_type            = typename variable_traits<VariableType>::index_type;
    static constexpr size_t nb_variables = variable_traits<VariableType>::nb_variables;
    using step_index_type                = typename step_traits<StepType>::index_type;
    static constexpr size_t nb_steps     = step_traits<StepType>::nb_steps;
    using view_type                      = typename SubgridType::template accessor_type<float_type>;

    SubgridMemoryManager(size_t nb_elements = 0, sc_MPI_Comm comm = sc_MPI_COMM_WORLD)
        : m_device_buffer(nb_elements * SubgridType::size, comm), m_device_volume(nb_elements, comm) {}
    ~SubgridMemoryManager()                                      = default;
    SubgridMemoryManager(SubgridMemoryManager&&)                 = default;   // move-only, like its storage
    SubgridMemoryManager& operator=(SubgridMemoryManager&&)      = default;
    SubgridMemoryManager(SubgridMemoryManager const&)            = delete;
    SubgridMemoryManager& operator=(SubgridMemoryManager const&) = delete;

    void set_variable(step_index_type step, variable_index_type variable,
                      thrust::device_vector<float_type> const& buffer) {
      m_device_buffer.copy(row(step, variable), buffer);
    }
    void set_variable(step_index_type step, variable_index_type variable,
                      thrust::host_vector<float_type> const& buffer) {
      m_device_buffer.copy(row(step, variable), buffer);
    }
    void set_variable(step_index_type step, variable_index_type variable, float_type* buffer) {
      m_device_buffer.copy(row(step, variable), buffer, m_device_buffer.size());
    }
    /// collective (resizes the shared volume vector).
    void set_volume(thrust::host_vector<float_type> const& buffer) { m_device_volume = buffer; }
    void set_volume(thrust::device_vector<float_type> const& buffer) { m_device_volume = buffer; }

    float_type*              get_own_volume() { return m_device_volume.get_own(); }
    float_type const*        get_own_volume() const { return m_device_volume.get_own(); }
    float_type* const*       get_all_volume() { return m_device_volume.get_all(); }
    float_type const* const* get_all_volume() const { return m_device_volume.get_all(); }

    [[nodiscard]] SubgridMemoryAccessorOwn<VariableType, SubgridType> get_own_variables(step_index_type step) {
      std::array<float_type*, nb_variables> a{};
      for (size_t k = 0; k < nb_variables; k++) a[k] = m_device_buffer.get_own(static_cast<int>(step * nb_variables + k));
      return SubgridMemoryAccessorOwn<VariableType, SubgridType>{a};
    }
    [[nodiscard]] SubgridMemoryAccessorAll<VariableType, SubgridType> get_all_variables(step_index_type step) {
      std::array<float_type* const*, nb_variables> a{};
      for (size_t k = 0; k < nb_variables; k++) a[k] = m_device_buffer.get_all(static_cast<int>(step * nb_variables + k));
      return SubgridMemoryAccessorAll<VariableType, SubgridType>{a};
    }
    [[nodiscard]] view_type get_own_variable(step_index_type step, variable_index_type variable) {
      return view_type{m_device_buffer.get_own(row(step, variable))};
    }
    [[nodiscard]] view_type const get_own_variable(step_index_type step, variable_index_type variable) const {
      return view_type{m_device_buffer.get_own(row(step, variable))};
    }

    /// new_size = number of ELEMENTS; collective.
    inline void resize(size_t new_size) { m_device_buffer.resize(new_size * SubgridType::size); }

   private:
    static int row(step_index_type s, variable_index_type v) {
      return static_cast<int>(s) * static_cast<int>(nb_variables) + static_cast<int>(v);
    }
    SharedDeviceVector<std::array<float_type, nb_variables * nb_steps>> m_device_buffer;
    SharedDeviceVector<float_type>                                      m_device_volume;
  };

}  // namespace t8gpu

#endif  // T8GPU_B200_MEMORY_SUBGRID_MEMORY_MANAGER_H
