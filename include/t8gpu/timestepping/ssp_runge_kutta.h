/// @file ssp_runge_kutta.h
/// @brief The six SSP-RK3 stage kernels, launchable exactly like the reference's
///        (t8gpu/timestepping/ssp_runge_kutta.h:18-116, .inl:30-221):
///
///   timestepping::SSP_3RK_step1<VariableList><<<blocks, 256>>>(prev, step1, fluxes, volume, dt, num_elements);
///   timestepping::subgrid::SSP_3RK_step1<VariableList, Subgrid<4,4,4>><<<num_elements, Subgrid::block_size>>>(...);
///
/// Arithmetic kept from the reference: the truncated literals 0.33333333333333 / 0.66666666666666 and the
/// left-to-right `c * dt / volume * flux` evaluation (.inl:23-25,43,66-68,91-93); the flux accumulators are zeroed.
/// The scale `c * dt / volume` is hoisted out of the variable loop (same value, one divide instead of nb_variables).
/// A solver that wants the stage fused with the flux evaluation uses <t8gpu/b200/fused.h> instead.
#ifndef T8GPU_B200_TIMESTEPPING_SSP_RUNGE_KUTTA_H
#define T8GPU_B200_TIMESTEPPING_SSP_RUNGE_KUTTA_H

#include <t8gpu/memory/memory_manager.h>
#include <t8gpu/memory/subgrid_memory_manager.h>

namespace t8gpu::timestepping {

  template<typename ft>
  struct rk_coeffs {
    static constexpr ft stage_2_1 = static_cast<ft>(0.75);
    static constexpr ft stage_2_2 = static_cast<ft>(0.25);
    static constexpr ft stage_2_3 = static_cast<ft>(0.25);
    static constexpr ft stage_3_1 = static_cast<ft>(0.33333333333333);
    static constexpr ft stage_3_2 = static_cast<ft>(0.66666666666666);
    static constexpr ft stage_3_3 = static_cast<ft>(0.66666666666666);
  };

  namespace detail {
    template<int STAGE, typename ft>
    __device__ __forceinline__ ft combine(ft prev, ft in, ft flux, ft scale) {
      if constexpr (STAGE == 1) return prev + scale * flux;
      else if constexpr (STAGE == 2) return rk_coeffs<ft>::stage_2_1 * prev + rk_coeffs<ft>::stage_2_2 * in + scale * flux;
      else return rk_coeffs<ft>::stage_3_1 * prev + rk_coeffs<ft>::stage_3_2 * in + scale * flux;
    }
    template<int STAGE, typename ft>
    __device__ __forceinline__ ft scale(ft delta_t, ft volume) {
      if constexpr (STAGE == 1) return delta_t / volume;
      else if constexpr (STAGE == 2) return rk_coeffs<ft>::stage_2_3 * delta_t / volume;
      else return rk_coeffs<ft>::stage_3_3 * delta_t / volume;
    }
    template<int STAGE, typename VariableType>
    __device__ __forceinline__ void element_stage(MemoryAccessorOwn<VariableType>& prev,
                                                  MemoryAccessorOwn<VariableType>& in,
                                                  MemoryAccessorOwn<VariableType>& out,
                                                  MemoryAccessorOwn<VariableType>& fluxes,
                                                  typename variable_traits<VariableType>::float_type const* volume,
                                                  typename variable_traits<VariableType>::float_type        delta_t,
                                                  int                                                       num_elements) {
      using ft    = typename variable_traits<VariableType>::float_type;
      int const i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= num_elements) return;
      ft const s = scale<STAGE, ft>(delta_t, volume[i]);
#pragma unroll
      for (size_t k = 0; k < variable_traits<VariableType>::nb_variables; k++) {
        ft const u    = STAGE == 1 ? ft{0} : in.get(k)[i];
        out.get(k)[i] = combine<STAGE, ft>(prev.get(k)[i], u, fluxes.get(k)[i], s);
        fluxes.get(k)[i] = ft{0};
      }
    }
    template<int STAGE, typename VariableType, typename SubgridType>
    __device__ __forceinline__ void subgrid_stage(SubgridMemoryAccessorOwn<VariableType, SubgridType>& prev,
                                                  SubgridMemoryAccessorOwn<VariableType, SubgridType>& in,
                                                  SubgridMemoryAccessorOwn<VariableType, SubgridType>& out,
                                                  SubgridMemoryAccessorOwn<VariableType, SubgridType>& fluxes,
                                                  typename variable_traits<VariableType>::float_type const* volumes,
                                                  typename variable_traits<VariableType>::float_type        delta_t) {
      using ft       = typename variable_traits<VariableType>::float_type;
      size_t const e = blockIdx.x;
      // one block per element, blockDim = Subgrid::block_size; the cell is addressed through its flat index
      int const    c = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
      ft const     v = volumes[e] / static_cast<ft>(SubgridType::size);
      ft const     s = scale<STAGE, ft>(delta_t, v);
#pragma unroll
      for (size_t l = 0; l < variable_traits<VariableType>::nb_variables; l++) {
        ft* const p = static_cast<ft*>(prev.get(l)) + e * SubgridType::size + c;
        ft* const o = static_cast<ft*>(out.get(l)) + e * SubgridType::size + c;
        ft* const f = static_cast<ft*>(fluxes.get(l)) + e * SubgridType::size + c;
        ft const  u = STAGE == 1 ? ft{0} : *(static_cast<ft*>(in.get(l)) + e * SubgridType::size + c);
        *o          = combine<STAGE, ft>(*p, u, *f, s);
        *f          = ft{0};
      }
    }
  }  // namespace detail

  /// step1 = prev + dt/vol * fluxes ; fluxes = 0
  template<typename VariableType>
  __global__ void SSP_3RK_step1(MemoryAccessorOwn<VariableType> prev, MemoryAccessorOwn<VariableType> step1,
                                MemoryAccessorOwn<VariableType> fluxes,
                                typename variable_traits<VariableType>::float_type const* __restrict__ volume,
                                typename variable_traits<VariableType>::float_type delta_t, int num_elements) {
    detail::element_stage<1, VariableType>(prev, prev, step1, fluxes, volume, delta_t, num_elements);
  }
  /// step2 = 3/4 prev + 1/4 step1 + 1/4 dt/vol * fluxes ; fluxes = 0
  template<typename VariableType>
  __global__ void SSP_3RK_step2(MemoryAccessorOwn<VariableType> prev, MemoryAccessorOwn<VariableType> step1,
                                MemoryAccessorOwn<VariableType> step2, MemoryAccessorOwn<VariableType> fluxes,
                                typename variable_traits<VariableType>::float_type const* __restrict__ volume,
                                typename variable_traits<VariableType>::float_type delta_t, int num_elements) {
    detail::element_stage<2, VariableType>(prev, step1, step2, fluxes, volume, delta_t, num_elements);
  }
  /// next = c31 prev + c32 step2 + c33 dt/vol * fluxes ; fluxes = 0
  template<typename VariableType>
  __global__ void SSP_3RK_step3(MemoryAccessorOwn<VariableType> prev, MemoryAccessorOwn<VariableType> step2,
                                MemoryAccessorOwn<VariableType> next, MemoryAccessorOwn<VariableType> fluxes,
                                typename variable_traits<VariableType>::float_type const* __restrict__ volume,
                                typename variable_traits<VariableType>::float_type delta_t, int num_elements) {
    detail::element_stage<3, VariableType>(prev, step2, next, fluxes, volume, delta_t, num_elements);
  }

  namespace subgrid {
    template<typename VariableType, typename SubgridType>
    __global__ void SSP_3RK_step1(SubgridMemoryAccessorOwn<VariableType, SubgridType> prev,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> step1,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> fluxes,
                                  typename variable_traits<VariableType>::float_type const* __restrict__ volumes,
                                  typename variable_traits<VariableType>::float_type delta_t) {
      detail::subgrid_stage<1, VariableType, SubgridType>(prev, prev, step1, fluxes, volumes, delta_t);
    }
    template<typename VariableType, typename SubgridType>
    __global__ void SSP_3RK_step2(SubgridMemoryAccessorOwn<VariableType, SubgridType> prev,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> step1,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> step2,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> fluxes,
                                  typename variable_traits<VariableType>::float_type const* __restrict__ volumes,
                                  typename variable_traits<VariableType>::float_type delta_t) {
      detail::subgrid_stage<2, VariableType, SubgridType>(prev, step1, step2, fluxes, volumes, delta_t);
    }
    template<typename VariableType, typename SubgridType>
    __global__ void SSP_3RK_step3(SubgridMemoryAccessorOwn<VariableType, SubgridType> prev,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> step2,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> next,
                                  SubgridMemoryAccessorOwn<VariableType, SubgridType> fluxes,
                                  typename variable_traits<VariableType>::float_type const* __restrict__ volumes,
                                  typename variable_traits<VariableType>::float_type delta_t) {
      detail::subgrid_stage<3, VariableType, SubgridType>(prev, step2, next, fluxes, volumes, delta_t);
    }
  }  // namespace subgrid

}  // namespace t8gpu::timestepping

#endif  // T8GPU_B200_TIMESTEPPING_SSP_RUNGE_KUTTA_H
