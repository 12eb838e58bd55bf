namespace gpb {
struct Prog {
  static constexpr int N_HP = 4, DIM = 1;
  static __device__ __forceinline__ double value(const double* __restrict__ h, const double* __restrict__ ih, const double* __restrict__ xi, const double* __restrict__ xj, int gi, int gj) {
    double r2 = 0.0;
    { const double t = xi[0] - xj[0]; r2 += t * t; }
    double l1 = 0.0;
    l1 += fabs(xi[0] - xj[0]);
    // node 0: leaf op 1 at hp[0]
    const double il0 = ih[0];
    const double q0 = r2 * (il0 * il0);
    const double k0_0 = exp(-0.5 * q0);
    const double v0 = k0_0;
    // node 1: leaf op 2 at hp[1]
    const double il1 = ih[1], ip1 = ih[2];
    const double u1 = 3.14159265358979323846 * (l1 / h[2]);
    const double s1 = gpb_sin(u1);
    const double sine1 = s1 * s1;
    const double il2_1 = il1 * il1;
    const double k0_1 = exp((-2.0 * sine1) * il2_1);
    const double v1 = k0_1;
    const double v2 = v0 + v1;
    // node 3: leaf op 3 at hp[3]
    double k0_3 = 0.0;
    k0_3 += (xi[0] - h[3]) * (xj[0] - h[3]);
    const double v3 = k0_3;
    const double v4 = v2 * v3;
    return v4;
  }
  static __device__ __forceinline__ void grad(const double* __restrict__ h, const double* __restrict__ ih, const double* __restrict__ xi, const double* __restrict__ xj, int gi, int gj, double w, double (&g)[N_HP + 1]) {
    double r2 = 0.0;
    { const double t = xi[0] - xj[0]; r2 += t * t; }
    double l1 = 0.0;
    l1 += fabs(xi[0] - xj[0]);
    // node 0: leaf op 1 at hp[0]
    const double il0 = ih[0];
    const double q0 = r2 * (il0 * il0);
    const double k0_0 = exp(-0.5 * q0);
    double d0_0 = k0_0 * q0 * il0;
    const double v0 = k0_0;
    // node 1: leaf op 2 at hp[1]
    const double il1 = ih[1], ip1 = ih[2];
    const double u1 = 3.14159265358979323846 * (l1 / h[2]);
    double s1, c1; gpb_sincos(u1, &s1, &c1);
    const double sine1 = s1 * s1;
    const double il2_1 = il1 * il1;
    const double k0_1 = exp((-2.0 * sine1) * il2_1);
    double d1_0 = k0_1 * (4.0 * sine1) * (il2_1 * il1);
    double d1_1 = k0_1 * (2.0 * 3.14159265358979323846 * l1 * (2.0 * s1 * c1)) * (il2_1 * (ip1 * ip1));
    const double v1 = k0_1;
    const double v2 = v0 + v1;
    // node 3: leaf op 3 at hp[3]
    double k0_3 = 0.0;
    k0_3 += (xi[0] - h[3]) * (xj[0] - h[3]);
    double d3_0 = 2.0 * h[3] - xi[0] - xj[0];
    const double v3 = k0_3;
    const double v4 = v2 * v3;
    const double a4 = w;
    const double a2 = a4 * v3;
    const double a3 = a4 * v2;
    g[3] += a3 * d3_0;
    const double a0 = a2;
    const double a1 = a2;
    g[1] += a1 * d1_0;
    g[2] += a1 * d1_1;
    g[0] += a0 * d0_0;
  }
};
}  // namespace gpb
extern "C" __global__ void __launch_bounds__(256, 1) gpb_spec_assemble(const GpbMat* __restrict__ mats, const int* __restrict__ which) {
  gpb::assemble_spec_body<gpb::Prog>(mats, which);
}
extern "C" __global__ void __launch_bounds__(256, 1) gpb_spec_grad(const GpbMat* __restrict__ mats, const int* __restrict__ which) {
  gpb::grad_spec_body<gpb::Prog>(mats, which);
}
