// double-storage, degree-3 instantiations of the persistent LSM sweep (lsm_resident_kernel.cuh).
#include "lsm_resident_kernel.cuh"

namespace optmc {

int launch_resident_f64_deg3(optmc_ctx* ctx, const ResPlan& p, ResArgs& a) {
  return launch_resident_shape<double, 3>(ctx, p, a);
}

}  // namespace optmc
