#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ double cvt_raw(float f) {   // value * 2^-896
    const unsigned u = __float_as_uint(f);
    unsigned long long w;
    asm("mul.wide.u32 %0, %1, 0x20000000;" : "=l"(w) : "r"(u & 0x7fffffffu));
    w |= (unsigned long long)(u & 0x80000000u) << 32;
    return __longlong_as_double((long long)w);
}
__device__ __forceinline__ double cvt_bias(float f) {
    const unsigned u = __float_as_uint(f);
    unsigned long long w;
    asm("mul.wide.u32 %0, %1, 0x20000000;" : "=l"(w) : "r"(u & 0x7fffffffu));
    const unsigned s = (u & 0x80000000u) | 0x38000000u;
    w += (unsigned long long)s << 32;
    return __longlong_as_double((long long)w);
}
__device__ __forceinline__ double rnd_a(double x) {     // predicate form
    unsigned lo = (unsigned)__double2loint(x), hi = (unsigned)__double2hiint(x);
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    b += (lo & 0x20000000u) ? 0x10000000ull : 0x0FFFFFFFull;
    b &= 0xFFFFFFFFE0000000ull;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double rnd_b(double x) {     // half away
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    b += 0x10000000ull;
    b &= 0xFFFFFFFFE0000000ull;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double rnd_c(double x) {     // asm carry chain
    unsigned lo = (unsigned)__double2loint(x), hi = (unsigned)__double2hiint(x);
    unsigned lo2, hi2;
    asm("{\n\t.reg .u32 t;\n\t"
        "shl.b32 t, %2, 2;\n\t"
        "add.cc.u32 t, t, 0x80000000;\n\t"
        "addc.cc.u32 %0, %2, 0x0FFFFFFF;\n\t"
        "addc.u32 %1, %3, 0;\n\t}"
        : "=r"(lo2), "=r"(hi2) : "r"(lo), "r"(hi));
    return __hiloint2double((int)hi2, (int)(lo2 & 0xE0000000u));
}
extern "C" __global__ void k_raw(const float* in, double* out) { int i = threadIdx.x; out[i] = cvt_raw(in[i]) * 3.0; }
extern "C" __global__ void k_bias(const float* in, double* out) { int i = threadIdx.x; out[i] = cvt_bias(in[i]) * 3.0; }
extern "C" __global__ void k_rnda(const double* in, double* out) { int i = threadIdx.x; out[i] = rnd_a(in[i] * 3.0) * 5.0; }
extern "C" __global__ void k_rndb(const double* in, double* out) { int i = threadIdx.x; out[i] = rnd_b(in[i] * 3.0) * 5.0; }
extern "C" __global__ void k_rndc(const double* in, double* out) { int i = threadIdx.x; out[i] = rnd_c(in[i] * 3.0) * 5.0; }
