/// @file shared_device_vector.h
/// @brief Device storage shared between the ranks of one node; same public interface as
///        t8gpu/memory/shared_device_vector.h:41-337.
///
/// Differences from the reference implementation (all behind the same interface):
///  - allocations go through the t8gpu_b200 C ABI (t8b200_shared_alloc / _open): zero-initialised (the reference leaves
///    a grown allocation uninitialised, which lets stale flux accumulators leak into the next step), and every row of
///    the SoA specialisation starts on a 128-byte boundary (capacity is a multiple of 32 elements) so rows can be
///    read with 128-bit loads and TMA bulk copies;
///  - with ranks on different GPUs of an NVSwitch node the peer pointers in the [row][rank] table are NVLink peer
///    mappings; with all ranks on one GPU (the reference's mode) they are plain IPC mappings.
/// The handle exchange stays an MPI_Allgather (the host environment of a t8gpu user has MPI through t8code).
#ifndef T8GPU_B200_MEMORY_SHARED_DEVICE_VECTOR_H
#define T8GPU_B200_MEMORY_SHARED_DEVICE_VECTOR_H

#include <sc.h>
#include <t8gpu/utils/cuda.h>
#include <t8gpu_b200.h>
#include <thrust/device_vector.h>
#include <thrust/host_vector.h>

#include <array>
#include <cassert>
#include <cstring>
#include <utility>
#include <vector>

namespace t8gpu {

  namespace detail {
    /// N rows of `capacity` elements of T in one allocation per rank, mapped on every rank.
    template<typename T>
    class SharedRows {
     public:
      SharedRows(size_t nrows, size_t size, sc_MPI_Comm comm) : m_comm{comm}, m_nrows{nrows} {
        MPI_Comm_size(m_comm, &m_nb_ranks);
        MPI_Comm_rank(m_comm, &m_rank);
        m_base.assign(m_nb_ranks, nullptr);
        m_cap.assign(m_nb_ranks, 0);
        m_rows.assign(m_nrows * m_nb_ranks, nullptr);
        reallocate_and_exchange(size, size);
      }
      ~SharedRows() { release(); }
      SharedRows(SharedRows const&)            = delete;
      SharedRows& operator=(SharedRows const&) = delete;
      SharedRows(SharedRows&& o) noexcept { steal(std::move(o)); }
      SharedRows& operator=(SharedRows&& o) noexcept {
        if (this != &o) {
          release();
          steal(std::move(o));
        }
        return *this;
      }

      /// collective; keeps the allocation when it is large enough, else grows to 1.5x (data is discarded, as in the
      /// reference: shared_device_vector.h:110-125).
      void resize(size_t new_size) {
        if (new_size <= m_cap[m_rank]) {
          m_size = new_size;
          exchange(false);
        } else {
          reallocate_and_exchange(new_size, new_size + new_size / 2);
        }
      }
      [[nodiscard]] size_t size() const { return m_size; }
      void                 clear() { m_size = 0; }
      [[nodiscard]] T*     own(size_t row) const { return m_rows[row * m_nb_ranks + m_rank]; }
      [[nodiscard]] T**    all(size_t row) { return thrust::raw_pointer_cast(m_device_rows.data()) + row * m_nb_ranks; }
      [[nodiscard]] T const* const* all(size_t row) const {
        return thrust::raw_pointer_cast(m_device_rows.data()) + row * m_nb_ranks;
      }
      [[nodiscard]] size_t capacity() const { return m_cap[m_rank]; }

     private:
      struct Wire {
        unsigned char handle[64];
        size_t        capacity;
        int           fresh;
      };
      /// The new allocation is made and published BEFORE the old one is freed, and the old one is freed only after every
      /// peer has closed its mapping of it (exchange() ends with a barrier when a rank published a new buffer): freeing an
      /// exported allocation that is still imported elsewhere is undefined behaviour (the reference frees first,
      /// shared_device_vector.inl:249-283).
      void reallocate_and_exchange(size_t size, size_t want_capacity) {
        T* const old   = m_base[m_rank];
        m_base[m_rank] = nullptr;
        m_size         = size;
        size_t cap     = (want_capacity + 31) / 32 * 32;  // 128-byte aligned rows for 4- and 8-byte T
        m_cap[m_rank]  = cap;
        std::memset(&m_wire, 0, sizeof(m_wire));
        if (cap > 0) {
          void* p = nullptr;
          T8GPU_CUDA_CHECK_ERROR(t8b200_shared_alloc(sizeof(T) * cap * m_nrows, &p, m_wire.handle));
          m_base[m_rank] = static_cast<T*>(p);
        }
        m_wire.capacity = cap;
        exchange(true);
        if (old) T8GPU_CUDA_CHECK_ERROR(t8b200_shared_free(old));
      }
      void exchange(bool fresh) {
        m_wire.fresh = fresh ? 1 : 0;
        std::vector<Wire> all(m_nb_ranks);
        all[m_rank] = m_wire;
        MPI_Allgather(MPI_IN_PLACE, 0, MPI_DATATYPE_NULL, all.data(), sizeof(Wire), MPI_BYTE, m_comm);
        bool any_fresh = false;
        for (int r = 0; r < m_nb_ranks; r++) {
          any_fresh = any_fresh || all[r].fresh;
          if (r != m_rank && all[r].fresh) {
            if (m_base[r]) T8GPU_CUDA_CHECK_ERROR(t8b200_shared_close(m_base[r]));
            m_base[r] = nullptr;
            m_cap[r]  = all[r].capacity;
            if (m_cap[r] > 0) {
              void* p = nullptr;
              T8GPU_CUDA_CHECK_ERROR(t8b200_shared_open(all[r].handle, &p));
              m_base[r] = static_cast<T*>(p);
            }
          }
          for (size_t k = 0; k < m_nrows; k++)
            m_rows[k * m_nb_ranks + r] = m_base[r] ? m_base[r] + k * m_cap[r] : nullptr;
        }
        m_device_rows = m_rows;
        if (any_fresh && m_nb_ranks > 1) MPI_Barrier(m_comm);   // every stale mapping is closed: old buffers may be freed
      }
      /// collective (destruction and move-assignment of the managers are, as in the reference): every rank first closes
      /// its mappings of the peers' buffers, then -- after a barrier -- frees its own.
      void release() {
        bool owns = false;
        for (int r = 0; r < m_nb_ranks; r++) {
          if (!m_base.empty() && m_base[r]) {
            if (r == m_rank) { owns = true; continue; }
            t8b200_shared_close(m_base[r]);
            m_base[r] = nullptr;
          }
        }
        if (m_nb_ranks > 1) MPI_Barrier(m_comm);
        if (owns) {
          t8b200_shared_free(m_base[m_rank]);
          m_base[m_rank] = nullptr;
        }
      }
      void steal(SharedRows&& o) {
        m_comm = o.m_comm; m_rank = o.m_rank; m_nb_ranks = o.m_nb_ranks; m_nrows = o.m_nrows; m_size = o.m_size;
        m_wire = o.m_wire;
        m_base = std::move(o.m_base); m_cap = std::move(o.m_cap); m_rows = std::move(o.m_rows);
        m_device_rows = std::move(o.m_device_rows);
        o.m_base.clear();
        o.m_nb_ranks = 0;
      }

      sc_MPI_Comm               m_comm{};
      int                       m_rank{0}, m_nb_ranks{0};
      size_t                    m_nrows{0}, m_size{0};
      Wire                      m_wire{};
      std::vector<T*>           m_base;
      std::vector<size_t>       m_cap;
      thrust::host_vector<T*>   m_rows;
      thrust::device_vector<T*> m_device_rows;
    };
  }  // namespace detail

  /// One array per rank, readable from every rank (shared_device_vector.h:41-163).
  template<typename T>
  class SharedDeviceVector {
   public:
    inline SharedDeviceVector(size_t size = 0, sc_MPI_Comm comm = sc_MPI_COMM_WORLD) : m_rows{1, size, comm} {}
    inline SharedDeviceVector(SharedDeviceVector&&)                 = default;
    inline SharedDeviceVector& operator=(SharedDeviceVector&&)      = default;
    inline SharedDeviceVector(SharedDeviceVector const&)            = delete;
    inline SharedDeviceVector& operator=(SharedDeviceVector const&) = delete;

    inline void resize(size_t new_size) { m_rows.resize(new_size); }
    /// collective: resizes to other.size() and copies.
    inline SharedDeviceVector<T> const& operator=(thrust::host_vector<T> const& other) {
      m_rows.resize(other.size());
      if (other.size())
        T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_rows.own(0), thrust::raw_pointer_cast(other.data()),
                                          sizeof(T) * other.size(), cudaMemcpyHostToDevice));
      return *this;
    }
    inline SharedDeviceVector<T> const& operator=(thrust::device_vector<T> const& other) {
      m_rows.resize(other.size());
      if (other.size())
        T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_rows.own(0), thrust::raw_pointer_cast(other.data()),
                                          sizeof(T) * other.size(), cudaMemcpyDeviceToDevice));
      return *this;
    }
    [[nodiscard]] inline size_t size() const { return m_rows.size(); }
    inline void                 clear() { m_rows.clear(); }
    [[nodiscard]] inline T*     get_own() { return m_rows.own(0); }
    [[nodiscard]] inline T**    get_all() { return m_rows.all(0); }
    [[nodiscard]] inline T const* get_own() const { return m_rows.own(0); }
    [[nodiscard]] inline T const* const* get_all() const { return m_rows.all(0); }

   private:
    detail::SharedRows<T> m_rows;
  };

  /// Struct-of-arrays specialisation: N arrays (rows) per rank in one allocation, row k at k * capacity
  /// (shared_device_vector.h:177-337).
  template<typename T, size_t N>
  class SharedDeviceVector<std::array<T, N>> {
   public:
    inline SharedDeviceVector(size_t size = 0, sc_MPI_Comm comm = sc_MPI_COMM_WORLD) : m_rows{N, size, comm} {}
    inline SharedDeviceVector(SharedDeviceVector&&)                 = default;
    inline SharedDeviceVector& operator=(SharedDeviceVector&&)      = default;
    inline SharedDeviceVector(SharedDeviceVector const&)            = delete;
    inline SharedDeviceVector& operator=(SharedDeviceVector const&) = delete;

    inline void resize(size_t new_size) { m_rows.resize(new_size); }
    inline void copy(size_t index, thrust::host_vector<T> const& vector) {
      assert(m_rows.size() <= vector.size());
      T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_rows.own(index), thrust::raw_pointer_cast(vector.data()),
                                        sizeof(T) * m_rows.size(), cudaMemcpyHostToDevice));
    }
    inline void copy(size_t index, thrust::device_vector<T> const& vector) {
      assert(m_rows.size() <= vector.size());
      T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_rows.own(index), thrust::raw_pointer_cast(vector.data()),
                                        sizeof(T) * m_rows.size(), cudaMemcpyDeviceToDevice));
    }
    inline void copy(size_t index, T const* buffer, size_t num_elements) {
      assert(num_elements <= m_rows.size());
      T8GPU_CUDA_CHECK_ERROR(cudaMemcpy(m_rows.own(index), buffer, sizeof(T) * num_elements, cudaMemcpyDeviceToDevice));
    }
    [[nodiscard]] inline size_t size() const { return m_rows.size(); }
    inline void                 clear() { m_rows.clear(); }
    [[nodiscard]] inline T*     get_own(int index) { return m_rows.own(index); }
    [[nodiscard]] inline T**    get_all(int index) { return m_rows.all(index); }
    [[nodiscard]] inline T const* get_own(int index) const { return m_rows.own(index); }
    [[nodiscard]] inline T const* const* get_all(int index) const { return m_rows.all(index); }
    /// stride between consecutive rows (t8gpu_b200 extension).
    [[nodiscard]] inline size_t capacity() const { return m_rows.capacity(); }

   private:
    detail::SharedRows<T> m_rows;
  };

}  // namespace t8gpu

#endif  // T8GPU_B200_MEMORY_SHARED_DEVICE_VECTOR_H
