// TEST-ONLY: compiles a user-style translation unit against the header mirror in include/t8gpu/ (with the t8code / sc /
// MPI declarations of oracle/ref_shim, as t8code itself is not installed here).  Exercises the API surface of
// SURVEY App. B the way the reference's example solvers use it.
#include <t8gpu/memory/memory_manager.h>
#include <t8gpu/memory/subgrid_memory_manager.h>
#include <t8gpu/timestepping/ssp_runge_kutta.h>
#include <t8gpu/utils/cuda.h>
#include <t8gpu/utils/meta.h>
#include <t8gpu/utils/profiling.h>

enum VariableList { Rho, Rho_v1, Rho_v2, Rho_v3, Rho_e, nb_variables };
enum StepList { Step0, Step1, Step2, Step3, Fluxes, nb_steps };

using namespace t8gpu;
using float_type = variable_traits<VariableList>::float_type;
using Sub        = Subgrid<4, 4, 4>;

__global__ void user_kernel(MemoryAccessorOwn<VariableList> own, MemoryAccessorAll<VariableList> all,
                            float_type const* volume, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  auto [rho, rho_e] = own.get(Rho, Rho_e);
  float_type* const* r_all = all.get(Rho);
  rho[i] = r_all[0][i] + rho_e[i] * volume[i];
}

__global__ void user_subgrid_kernel(SubgridMemoryAccessorOwn<VariableList, Sub> own,
                                    SubgridMemoryAccessorAll<VariableList, Sub> all) {
  int e = blockIdx.x, i = threadIdx.x, j = threadIdx.y, k = threadIdx.z;
  auto rho = own.get(Rho);
  rho(e, i, j, k) = all.get(0, Rho)(e, i, j, k) + float_type(Sub::flat_index(i, j, k));
}

int user_code(sc_MPI_Comm comm) {
  static_assert(Sub::rank == 3 && Sub::size == 64 && Sub::extent<0> == 4 && Sub::stride<1> == 4, "Subgrid");
  static_assert(meta::log2_v<4> == 2 && meta::all_same_v<int, int>, "meta");
  MemoryManager<VariableList, StepList> mem(1000, comm);
  thrust::host_vector<float_type> h(1000, float_type(1));
  mem.set_variable(Step0, Rho, h);
  mem.set_volume(h);
  user_kernel<<<4, 256>>>(mem.get_own_variables(Step0), mem.get_all_variables(Step0), mem.get_own_volume(), 1000);
  T8GPU_CUDA_CHECK_LAST_ERROR();
  timestepping::SSP_3RK_step1<VariableList><<<4, 256>>>(mem.get_own_variables(Step0), mem.get_own_variables(Step1),
                                                        mem.get_own_variables(Fluxes), mem.get_own_volume(),
                                                        float_type(0.1), 1000);
  timestepping::SSP_3RK_step2<VariableList><<<4, 256>>>(mem.get_own_variables(Step0), mem.get_own_variables(Step1),
                                                        mem.get_own_variables(Step2), mem.get_own_variables(Fluxes),
                                                        mem.get_own_volume(), float_type(0.1), 1000);
  timestepping::SSP_3RK_step3<VariableList><<<4, 256>>>(mem.get_own_variables(Step0), mem.get_own_variables(Step2),
                                                        mem.get_own_variables(Step3), mem.get_own_variables(Fluxes),
                                                        mem.get_own_volume(), float_type(0.1), 1000);
  mem.resize(2000);
  float_type* const*       tables = mem.get_all_variables(Step1).data()[0];
  float_type* const*       own    = mem.get_own_variables(Step1).data();
  (void)tables; (void)own;

  SubgridMemoryManager<VariableList, StepList, Sub> smem(10, comm);
  thrust::host_vector<float_type> hs(10 * Sub::size, float_type(1)), hv(10, float_type(1));
  smem.set_variable(Step0, Rho, hs);
  smem.set_volume(hv);
  user_subgrid_kernel<<<10, Sub::block_size>>>(smem.get_own_variables(Step0), smem.get_all_variables(Step0));
  timestepping::subgrid::SSP_3RK_step1<VariableList, Sub><<<10, Sub::block_size>>>(
      smem.get_own_variables(Step0), smem.get_own_variables(Step1), smem.get_own_variables(Fluxes),
      smem.get_own_volume(), float_type(0.1));
  T8GPU_CUDA_CHECK_ERROR(cudaDeviceSynchronize());
  return 0;
}
