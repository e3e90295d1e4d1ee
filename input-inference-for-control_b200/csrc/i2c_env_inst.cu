// One translation unit per environment: compiled with -DI2C_ENV_ID=<enum i2c_env> (see __graft_entry__.build()).
#include "i2c_kernels.cuh"
#include "i2c_scan.cuh"
#include "i2c_group.cuh"

#ifndef I2C_ENV_ID
#error "compile with -DI2C_ENV_ID=<0..6>"
#endif

namespace i2c {

#if I2C_ENV_ID == 0
using EnvT = EnvLinear;
#elif I2C_ENV_ID == 1
using EnvT = EnvLinearMinEnergy;
#elif I2C_ENV_ID == 2
using EnvT = EnvPendulum;
#elif I2C_ENV_ID == 3
using EnvT = EnvPendulumActReg;
#elif I2C_ENV_ID == 4
using EnvT = EnvCartpole;
#elif I2C_ENV_ID == 5
using EnvT = EnvDoubleCartpole;
#elif I2C_ENV_ID == 6
using EnvT = EnvQuadrotor;
#endif

#define I2C_CAT2(a, b) a##b
#define I2C_CAT(a, b) I2C_CAT2(a, b)

int I2C_CAT(launch_em_env, I2C_ENV_ID)(const KParams& p, void* stream) { return launch_em_t<EnvT>(p, (cudaStream_t)stream); }
int I2C_CAT(launch_quad_env, I2C_ENV_ID)(int fn, const QuadArgs& a, void* stream) {
  return launch_quad_t<EnvT>(fn, a, (cudaStream_t)stream);
}
int I2C_CAT(launch_ckf_env, I2C_ENV_ID)(const CkfArgs& a, void* stream) { return launch_ckf_t<EnvT>(a, (cudaStream_t)stream); }
int I2C_CAT(launch_rollout_env, I2C_ENV_ID)(const RolloutArgs& a, void* stream) {
  return launch_rollout_t<EnvT>(a, (cudaStream_t)stream);
}

int I2C_CAT(launch_scan_env, I2C_ENV_ID)(int stage, const KParams& p, const ScanArgs& a, void* stream) {
  return launch_scan_t<EnvT>(stage, p, a, (cudaStream_t)stream);
}

}  // namespace i2c
