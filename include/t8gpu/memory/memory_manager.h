/// @file memory_manager.h
/// @brief MemoryManager and the MemoryAccessor{Own,All} kernel-argument PODs; source-compatible with
///        t8gpu/memory/memory_manager.h:24-461 (same names, same get() overloads incl. the structured-binding form).
///
/// float_type: the reference hard-codes `float` (memory_manager.h:29,39).  Here it is `T8GPU_FLOAT_TYPE`
/// (default float), so `-DT8GPU_FLOAT_TYPE=double` gives the fp64 configuration without editing the library.
#ifndef T8GPU_B200_MEMORY_MEMORY_MANAGER_H
#define T8GPU_B200_MEMORY_MEMORY_MANAGER_H

#include <t8gpu/memory/shared_device_vector.h>
#include <t8gpu/utils/meta.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <array>
#include <tuple>
#include <type_traits>

#ifndef T8GPU_FLOAT_TYPE
#define T8GPU_FLOAT_TYPE float
#endif

namespace t8gpu {

  template<class VariableList, typename = void>
  struct variable_traits {};

  /// enum VariableList { ..., nb_variables }
  template<class VariableType>
  struct variable_traits<VariableType, std::enable_if_t<std::is_enum_v<VariableType>>> {
    using float_type                     = T8GPU_FLOAT_TYPE;
    using index_type                     = VariableType;
    static constexpr size_t nb_variables = VariableType::nb_variables;
  };

  template<class StepList, typename = void>
  struct step_traits {};

  /// enum StepList { ..., nb_steps }
  template<class StepType>
  struct step_traits<StepType, std::enable_if_t<std::is_enum_v<StepType>>> {
    using float_type                 = T8GPU_FLOAT_TYPE;
    using index_type                 = StepType;
    static constexpr size_t nb_steps = StepType::nb_steps;
  };

  template<typename VariableType, typename StepType>
  class MemoryManager;
  template<typename VariableType, typename StepType, typename SubgridType>
  class SubgridMemoryManager;
  template<typename VariableType, typename StepType, size_t dim>
  class MeshManager;
  template<typename VariableType, typename StepType, typename SubgridType>
  class SubgridMeshManager;

  namespace detail {
    /// shared implementation of the two accessors: Ptr = float_type* (own) or float_type* const* (all ranks)
    template<typename VariableType, typename Ptr, typename ConstPtr>
    class AccessorBase {
     public:
      using variable_index_type            = typename variable_traits<VariableType>::index_type;
      using float_type                     = typename variable_traits<VariableType>::float_type;
      constexpr static size_t nb_variables = variable_traits<VariableType>::nb_variables;

      AccessorBase(AccessorBase const&)            = default;
      AccessorBase& operator=(AccessorBase const&) = default;

      template<typename T>
      [[nodiscard]] __device__ __host__ inline std::enable_if_t<
          meta::is_explicitly_convertible_to_v<T, variable_index_type>, Ptr>
      get(T i) {
        return m_pointers[static_cast<variable_index_type>(i)];
      }
      template<typename T>
      [[nodiscard]] __device__ __host__ inline std::enable_if_t<
          meta::is_explicitly_convertible_to_v<T, variable_index_type>, ConstPtr>
      get(T i) const {
        return m_pointers[static_cast<variable_index_type>(i)];
      }
      /// auto [a, b] = accessor.get(A, B);
      template<typename... Ts>
      [[nodiscard]] __device__ __host__ inline std::enable_if_t<
          (sizeof...(Ts) > 1) &&
              meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
              meta::all_same_v<Ts...>,
          std::array<Ptr, sizeof...(Ts)>>
      get(Ts... is) {
        return {get(static_cast<variable_index_type>(is))...};
      }
      template<typename... Ts>
      [[nodiscard]] __device__ __host__ inline std::enable_if_t<
          (sizeof...(Ts) > 1) &&
              meta::is_explicitly_convertible_to_v<std::tuple_element_t<0, std::tuple<Ts...>>, variable_index_type> &&
              meta::all_same_v<Ts...>,
          std::array<ConstPtr, sizeof...(Ts)>>
      get(Ts... is) const {
        return {get(static_cast<variable_index_type>(is))...};
      }

      /// Host-side view of the pointer array, as the C ABI (t8gpu_b200.h) takes it.  The reference keeps the array
      /// private so that the implementation may change (memory_manager.h:172-185); this is the one addition.
      [[nodiscard]] __host__ Ptr const* data() const { return m_pointers.data(); }

     protected:
      std::array<Ptr, nb_variables> m_pointers;
      // (constrained so that it never competes with the copy constructor for a non-const accessor)
      template<typename Container, typename = std::enable_if_t<!std::is_base_of_v<AccessorBase, std::decay_t<Container>>>>
      AccessorBase(Container&& array) : m_pointers(std::forward<Container>(array)) {}
    };
  }  // namespace detail

  /// Variables of the elements owned by this rank: `accessor.get(Rho)[i]` (memory_manager.h:87-186).
  template<typename VariableType>
  class MemoryAccessorOwn
      : public detail::AccessorBase<VariableType, typename variable_traits<VariableType>::float_type*,
                                    typename variable_traits<VariableType>::float_type const*> {
    using Base = detail::AccessorBase<VariableType, typename variable_traits<VariableType>::float_type*,
                                      typename variable_traits<VariableType>::float_type const*>;
    template<typename VT, typename ST>
    friend class MemoryManager;
    template<typename VT, typename ST, size_t dim_>
    friend class MeshManager;
    template<typename VT, typename ST, typename SubgridType>
    friend class SubgridMeshManager;

   public:
    MemoryAccessorOwn(MemoryAccessorOwn const&)            = default;
    MemoryAccessorOwn& operator=(MemoryAccessorOwn const&) = default;

   private:
    template<typename Container,
             typename = std::enable_if_t<!std::is_same_v<std::decay_t<Container>, MemoryAccessorOwn>>>
    MemoryAccessorOwn(Container&& array) : Base(std::forward<Container>(array)) {}
  };

  /// Variables of every rank: `accessor.get(Rho)[rank][i]` (memory_manager.h:216-313).
  template<typename VariableType>
  class MemoryAccessorAll
      : public detail::AccessorBase<VariableType, typename variable_traits<VariableType>::float_type* const*,
                                    typename variable_traits<VariableType>::float_type const* const*> {
    using Base = detail::AccessorBase<VariableType, typename variable_traits<VariableType>::float_type* const*,
                                      typename variable_traits<VariableType>::float_type const* const*>;
    template<typename VT, typename ST>
    friend class MemoryManager;
    template<typename VT, typename ST, size_t dim_>
    friend class MeshManager;

   public:
    MemoryAccessorAll(MemoryAccessorAll const&)            = default;
    MemoryAccessorAll& operator=(MemoryAccessorAll const&) = default;

   private:
    template<typename Container,
             typename = std::enable_if_t<!std::is_same_v<std::decay_t<Container>, MemoryAccessorAll>>>
    MemoryAccessorAll(Container&& array) : Base(std::forward<Container>(array)) {}
  };

  /// Owns the device storage of nb_variables x nb_steps arrays + the volume (memory_manager.h:326-461).
  /// Row index = step * nb_variables + variable; the volume is row nb_steps * nb_variables.
  template<typename VariableType, typename StepType>
  class MemoryManager {
   public:
    using float_type                     = typename variable_traits<VariableType>::float_type;
    using variable_index_type            = typename variable_traits<VariableType>::index_type;
    static constexpr size_t nb_variables = variable_traits<VariableType>::nb_variables;
    using step_index_type                = typename step_traits<StepType>::index_type;
    static constexpr size_t nb_steps     = step_traits<StepType>::nb_steps;

    MemoryManager(size_t nb_elements = 0, sc_MPI_Comm comm = sc_MPI_COMM_WORLD) : m_device_buffer(nb_elements, comm) {}
    ~MemoryManager()                               = default;
    MemoryManager(MemoryManager&&)                 = default;   // move-only, like the storage it owns
    MemoryManager& operator=(MemoryManager&&)      = default;
    MemoryManager(MemoryManager const&)            = delete;
    MemoryManager& operator=(MemoryManager const&) = delete;

    void set_variable(step_index_type step, variable_index_type variable,
                      thrust::device_vector<float_type> const& buffer) {
      m_device_buffer.copy(row(step, variable), buffer);
    }
    void set_variable(step_index_type step, variable_index_type variable,
                      thrust::host_vector<float_type> const& buffer) {
      m_device_buffer.copy(row(step, variable), buffer);
    }
    /// buffer: device pointer to size() elements.
    void set_variable(step_index_type step, variable_index_type variable, float_type* buffer) {
      m_device_buffer.copy(row(step, variable), buffer, m_device_buffer.size());
    }
    void set_volume(thrust::host_vector<float_type> const& buffer) { m_device_buffer.copy(volume_row, buffer); }
    void set_volume(thrust::device_vector<float_type> const& buffer) { m_device_buffer.copy(volume_row, buffer); }
    void set_volume(float_type* buffer) { m_device_buffer.copy(volume_row, buffer, m_device_buffer.size()); }

    float_type*              get_own_volume() { return m_device_buffer.get_own(volume_row); }
    float_type const*        get_own_volume() const { return m_device_buffer.get_own(volume_row); }
    float_type* const*       get_all_volume() { return m_device_buffer.get_all(volume_row); }
    float_type const* const* get_all_volume() const { return m_device_buffer.get_all(volume_row); }

    [[nodiscard]] MemoryAccessorOwn<VariableType> get_own_variables(step_index_type step) {
      std::array<float_type*, nb_variables> a{};
      for (size_t k = 0; k < nb_variables; k++) a[k] = m_device_buffer.get_own(static_cast<int>(step * nb_variables + k));
      return MemoryAccessorOwn<VariableType>{a};
    }
    [[nodiscard]] MemoryAccessorAll<VariableType> get_all_variables(step_index_type step) {
      std::array<float_type* const*, nb_variables> a{};
      for (size_t k = 0; k < nb_variables; k++) a[k] = m_device_buffer.get_all(static_cast<int>(step * nb_variables + k));
      return MemoryAccessorAll<VariableType>{a};
    }
    [[nodiscard]] float_type* get_own_variable(step_index_type step, variable_index_type variable) {
      return m_device_buffer.get_own(row(step, variable));
    }
    [[nodiscard]] float_type const* get_own_variable(step_index_type step, variable_index_type variable) const {
      return m_device_buffer.get_own(row(step, variable));
    }

    /// collective over the communicator; discards the data when the allocation has to grow.
    inline void resize(size_t new_size) { m_device_buffer.resize(new_size); }

   private:
    static constexpr int volume_row = static_cast<int>(nb_steps * nb_variables);
    static int           row(step_index_type s, variable_index_type v) {
      return static_cast<int>(s) * static_cast<int>(nb_variables) + static_cast<int>(v);
    }
    SharedDeviceVector<std::array<float_type, nb_variables * nb_steps + 1>> m_device_buffer;
  };

}  // namespace t8gpu

#endif  // T8GPU_B200_MEMORY_MEMORY_MANAGER_H
