// cude_math.cuh — FP64 elementary functions used by the kernels.
// Kept behind m_* names so that hand-tuned versions can replace the CUDA library ones
// without touching the integrator.
#pragma once
#ifndef CUDE_HOST_EMU
#include <cuda_runtime.h>
#endif

namespace cude {

__device__ __forceinline__ double m_exp(double x) { return exp(x); }
__device__ __forceinline__ double m_tanh(double x) { return tanh(x); }
// softplus(x) = log(1 + exp(x)) — the naive form of reference src/neural-network.jl:13-15
__device__ __forceinline__ double m_softplus(double x) { return log(1.0 + exp(x)); }
// d softplus / dx
__device__ __forceinline__ double m_sigmoid(double x) { return 1.0 / (1.0 + exp(-x)); }
__device__ __forceinline__ double m_pow(double x, double y) { return pow(x, y); }
__device__ __forceinline__ double m_log10(double x) { return log10(x); }
__device__ __forceinline__ double m_pow10(double x) { return pow(10.0, x); }

}  // namespace cude
