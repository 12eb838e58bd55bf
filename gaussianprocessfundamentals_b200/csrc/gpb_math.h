// sin / cos for the periodic kernel (BaseKernels.py:446-453) without the library's large-argument slow path.
//
// CUDA's sin / sincos switch to a Payne-Hanek reduction for |a| > 105615 whose scratch array lives in local memory; it
// is the only local-memory access of the covariance kernels and is dead code here: the argument is pi * |x - x'| / p.
// This version reduces with a three-term Cody-Waite split of pi/2 under FMA (exact to well below an ulp for
// |a| < 2^31, far beyond where a double still resolves the phase) and evaluates the fdlibm kernels (< 1 ulp), so the
// generated kernels contain no LDL / STL at all.  Plain C++ (host + device): the host test harness and NVRTC compile it.
#pragma once
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define GPB_MHD __host__ __device__ __forceinline__
#else
#include <math.h>
#define GPB_MHD inline
#endif

GPB_MHD void gpb_sincos(double a, double* sp, double* cp) {
  const double q = rint(a * 0.63661977236758138);             // a * 2 / pi
  double r = fma(-q, 1.5707963267948966, a);                  // pi/2 = hi + mid + lo
  r = fma(-q, 6.123233995736766e-17, r);
  r = fma(-q, -1.4973849048591698e-33, r);
  const double z = r * r;
  // sin kernel on [-pi/4, pi/4]
  double ps = 1.58969099521155010221e-10;
  ps = fma(ps, z, -2.50507602534068634195e-08);
  ps = fma(ps, z, 2.75573137070700676789e-06);
  ps = fma(ps, z, -1.98412698298579493134e-04);
  ps = fma(ps, z, 8.33333333332248946124e-03);
  const double rs = r + (z * r) * fma(z, ps, -1.66666666666666324348e-01);
  // cos kernel
  double pc = -1.13596475577881948265e-11;
  pc = fma(pc, z, 2.08757232129817482790e-09);
  pc = fma(pc, z, -2.75573143513906633035e-07);
  pc = fma(pc, z, 2.48015872894767294178e-05);
  pc = fma(pc, z, -1.38888888888741095749e-03);
  pc = fma(pc, z, 4.16666666666666019037e-02);
  const double hz = 0.5 * z, w = 1.0 - hz;
  const double rc = w + (((1.0 - w) - hz) + z * (z * pc));
  const int iq = (int)(long long)q;
  const bool swap = (iq & 1) != 0;
  double s = swap ? rc : rs, c = swap ? rs : rc;
  if (iq & 2) { s = -s; c = -c; }
  if (swap) c = -c;
  *sp = s;
  *cp = c;
}

GPB_MHD double gpb_sin(double a) {
  double s, c;
  gpb_sincos(a, &s, &c);
  return s;
}
