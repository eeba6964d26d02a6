// idelucs_b200 — internal helpers shared by the .cu translation units.
#pragma once
#include <cuda_runtime.h>

#include "../../include/idelucs_b200.h"

namespace idl {
// records the thread-local error text returned by idl_last_error(); returns `code`
int set_error(int code, const char* fmt, const char* a = "", long long b = 0);
// counts the kernels this library has launched (idl_launch_count)
void note_launch();
// IIC loss for C <= 16 in one ordinary CTA (train_ops.cu)
constexpr int IID_SMALL_MAXC = 16;
int iid_loss_small_launch(const float* d_z1, const float* d_z2, int B, int C, float lamb, float eps, float* d_loss, float* d_joint, float* d_dz1,
                          float* d_dz2, void* stream, float gscale, float loss_w, const float* d_add, float add_w);
}  // namespace idl

#define IDL_CUDA_CHECK(expr)                                                                             \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            return ::idl::set_error(IDL_ECUDA, "%s (cuda error %lld)", cudaGetErrorString(_e), (long long)_e); \
    } while (0)
