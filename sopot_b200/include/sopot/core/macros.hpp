// SOPOT_HD marks what a component system needs on the device.  Compiled by g++ the macro is empty and the headers are
// plain C++20; compiled by nvcc (-std=c++20 --expt-relaxed-constexpr) every annotated function also exists as
// device code, which is what lets sopot/b200/generic_ensemble.cuh instantiate the RK4 ensemble kernel for a
// user-defined component pack inside the user's own translation unit.  SOPOT_UNROLL asks nvcc to unroll a constant-bound
// loop completely so that small per-instance arrays are indexed statically and stay in registers.
#pragma once
#ifdef __CUDACC__
#define SOPOT_HD __host__ __device__
#define SOPOT_B200_DEVICE_COMPILER 1
#define SOPOT_UNROLL _Pragma("unroll")
#else
#define SOPOT_HD
#define SOPOT_B200_DEVICE_COMPILER 0
#define SOPOT_UNROLL
#endif
