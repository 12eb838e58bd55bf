namespace gpb {
struct Prog {
  static constexpr int N_HP = 24, DIM = 1;
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
    // node 1: leaf op 3 at hp[1]
    double k0_1 = 0.0;
    k0_1 += (xi[0] - h[1]) * (xj[0] - h[1]);
    const double v1 = k0_1;
    const double v2 = v0 + v1;
    // node 3: leaf op 1 at hp[2]
    const double il3 = ih[2];
    const double q3 = r2 * (il3 * il3);
    const double k0_3 = exp(-0.5 * q3);
    const double v3 = k0_3;
    const double v4 = v2 + v3;
    // node 5: leaf op 1 at hp[3]
    const double il5 = ih[3];
    const double q5 = r2 * (il5 * il5);
    const double k0_5 = exp(-0.5 * q5);
    const double v5 = k0_5;
    // node 6: leaf op 1 at hp[4]
    const double il6 = ih[4];
    const double q6 = r2 * (il6 * il6);
    const double k0_6 = exp(-0.5 * q6);
    const double v6 = k0_6;
    const double v7 = v5 + v6;
    // node 8: leaf op 3 at hp[5]
    double k0_8 = 0.0;
    k0_8 += (xi[0] - h[5]) * (xj[0] - h[5]);
    const double v8 = k0_8;
    const double v9 = v7 + v8;
    const double v10 = v4 + v9;
    // node 11: leaf op 3 at hp[6]
    double k0_11 = 0.0;
    k0_11 += (xi[0] - h[6]) * (xj[0] - h[6]);
    const double v11 = k0_11;
    // node 12: leaf op 3 at hp[7]
    double k0_12 = 0.0;
    k0_12 += (xi[0] - h[7]) * (xj[0] - h[7]);
    const double v12 = k0_12;
    const double v13 = v11 * v12;
    const double v14 = v10 + v13;
    // node 15: leaf op 1 at hp[8]
    const double il15 = ih[8];
    const double q15 = r2 * (il15 * il15);
    const double k0_15 = exp(-0.5 * q15);
    const double v15 = k0_15;
    // node 16: leaf op 1 at hp[9]
    const double il16 = ih[9];
    const double q16 = r2 * (il16 * il16);
    const double k0_16 = exp(-0.5 * q16);
    const double v16 = k0_16;
    const double v17 = v15 + v16;
    // node 18: leaf op 3 at hp[10]
    double k0_18 = 0.0;
    k0_18 += (xi[0] - h[10]) * (xj[0] - h[10]);
    const double v18 = k0_18;
    const double v19 = v17 + v18;
    // node 20: leaf op 1 at hp[11]
    const double il20 = ih[11];
    const double q20 = r2 * (il20 * il20);
    const double k0_20 = exp(-0.5 * q20);
    const double v20 = k0_20;
    // node 21: leaf op 2 at hp[12]
    const double il21 = ih[12], ip21 = ih[13];
    const double u21 = 3.14159265358979323846 * (l1 / h[13]);
    const double s21 = gpb_sin(u21);
    const double sine21 = s21 * s21;
    const double il2_21 = il21 * il21;
    const double k0_21 = exp((-2.0 * sine21) * il2_21);
    const double v21 = k0_21;
    const double v22 = v20 + v21;
    // node 23: leaf op 1 at hp[14]
    const double il23 = ih[14];
    const double q23 = r2 * (il23 * il23);
    const double k0_23 = exp(-0.5 * q23);
    const double v23 = k0_23;
    const double v24 = v22 + v23;
    const double v25 = v19 + v24;
    const double v26 = v14 * v25;
    // node 27: leaf op 2 at hp[15]
    const double il27 = ih[15], ip27 = ih[16];
    const double u27 = 3.14159265358979323846 * (l1 / h[16]);
    const double s27 = gpb_sin(u27);
    const double sine27 = s27 * s27;
    const double il2_27 = il27 * il27;
    const double k0_27 = exp((-2.0 * sine27) * il2_27);
    const double v27 = k0_27;
    // node 28: leaf op 3 at hp[17]
    double k0_28 = 0.0;
    k0_28 += (xi[0] - h[17]) * (xj[0] - h[17]);
    const double v28 = k0_28;
    const double v29 = v27 + v28;
    // node 30: leaf op 3 at hp[18]
    double k0_30 = 0.0;
    k0_30 += (xi[0] - h[18]) * (xj[0] - h[18]);
    const double v30 = k0_30;
    const double v31 = v29 + v30;
    // node 32: leaf op 3 at hp[19]
    double k0_32 = 0.0;
    k0_32 += (xi[0] - h[19]) * (xj[0] - h[19]);
    const double v32 = k0_32;
    // node 33: leaf op 3 at hp[20]
    double k0_33 = 0.0;
    k0_33 += (xi[0] - h[20]) * (xj[0] - h[20]);
    const double v33 = k0_33;
    const double v34 = v32 * v33;
    // node 35: leaf op 1 at hp[21]
    const double il35 = ih[21];
    const double q35 = r2 * (il35 * il35);
    const double k0_35 = exp(-0.5 * q35);
    const double v35 = k0_35;
    const double v36 = v34 * v35;
    const double v37 = v31 * v36;
    // node 38: leaf op 3 at hp[22]
    double k0_38 = 0.0;
    k0_38 += (xi[0] - h[22]) * (xj[0] - h[22]);
    const double v38 = k0_38;
    // node 39: leaf op 3 at hp[23]
    double k0_39 = 0.0;
    k0_39 += (xi[0] - h[23]) * (xj[0] - h[23]);
    const double v39 = k0_39;
    const double v40 = v38 + v39;
    const double v41 = v37 * v40;
    const double v42 = v26 * v41;
    return v42;
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
    // node 1: leaf op 3 at hp[1]
    double k0_1 = 0.0;
    k0_1 += (xi[0] - h[1]) * (xj[0] - h[1]);
    double d1_0 = 2.0 * h[1] - xi[0] - xj[0];
    const double v1 = k0_1;
    const double v2 = v0 + v1;
    // node 3: leaf op 1 at hp[2]
    const double il3 = ih[2];
    const double q3 = r2 * (il3 * il3);
    const double k0_3 = exp(-0.5 * q3);
    double d3_0 = k0_3 * q3 * il3;
    const double v3 = k0_3;
    const double v4 = v2 + v3;
    // node 5: leaf op 1 at hp[3]
    const double il5 = ih[3];
    const double q5 = r2 * (il5 * il5);
    const double k0_5 = exp(-0.5 * q5);
    double d5_0 = k0_5 * q5 * il5;
    const double v5 = k0_5;
    // node 6: leaf op 1 at hp[4]
    const double il6 = ih[4];
    const double q6 = r2 * (il6 * il6);
    const double k0_6 = exp(-0.5 * q6);
    double d6_0 = k0_6 * q6 * il6;
    const double v6 = k0_6;
    const double v7 = v5 + v6;
    // node 8: leaf op 3 at hp[5]
    double k0_8 = 0.0;
    k0_8 += (xi[0] - h[5]) * (xj[0] - h[5]);
    double d8_0 = 2.0 * h[5] - xi[0] - xj[0];
    const double v8 = k0_8;
    const double v9 = v7 + v8;
    const double v10 = v4 + v9;
    // node 11: leaf op 3 at hp[6]
    double k0_11 = 0.0;
    k0_11 += (xi[0] - h[6]) * (xj[0] - h[6]);
    double d11_0 = 2.0 * h[6] - xi[0] - xj[0];
    const double v11 = k0_11;
    // node 12: leaf op 3 at hp[7]
    double k0_12 = 0.0;
    k0_12 += (xi[0] - h[7]) * (xj[0] - h[7]);
    double d12_0 = 2.0 * h[7] - xi[0] - xj[0];
    const double v12 = k0_12;
    const double v13 = v11 * v12;
    const double v14 = v10 + v13;
    // node 15: leaf op 1 at hp[8]
    const double il15 = ih[8];
    const double q15 = r2 * (il15 * il15);
    const double k0_15 = exp(-0.5 * q15);
    double d15_0 = k0_15 * q15 * il15;
    const double v15 = k0_15;
    // node 16: leaf op 1 at hp[9]
    const double il16 = ih[9];
    const double q16 = r2 * (il16 * il16);
    const double k0_16 = exp(-0.5 * q16);
    double d16_0 = k0_16 * q16 * il16;
    const double v16 = k0_16;
    const double v17 = v15 + v16;
    // node 18: leaf op 3 at hp[10]
    double k0_18 = 0.0;
    k0_18 += (xi[0] - h[10]) * (xj[0] - h[10]);
    double d18_0 = 2.0 * h[10] - xi[0] - xj[0];
    const double v18 = k0_18;
    const double v19 = v17 + v18;
    // node 20: leaf op 1 at hp[11]
    const double il20 = ih[11];
    const double q20 = r2 * (il20 * il20);
    const double k0_20 = exp(-0.5 * q20);
    double d20_0 = k0_20 * q20 * il20;
    const double v20 = k0_20;
    // node 21: leaf op 2 at hp[12]
    const double il21 = ih[12], ip21 = ih[13];
    const double u21 = 3.14159265358979323846 * (l1 / h[13]);
    double s21, c21; gpb_sincos(u21, &s21, &c21);
    const double sine21 = s21 * s21;
    const double il2_21 = il21 * il21;
    const double k0_21 = exp((-2.0 * sine21) * il2_21);
    double d21_0 = k0_21 * (4.0 * sine21) * (il2_21 * il21);
    double d21_1 = k0_21 * (2.0 * 3.14159265358979323846 * l1 * (2.0 * s21 * c21)) * (il2_21 * (ip21 * ip21));
    const double v21 = k0_21;
    const double v22 = v20 + v21;
    // node 23: leaf op 1 at hp[14]
    const double il23 = ih[14];
    const double q23 = r2 * (il23 * il23);
    const double k0_23 = exp(-0.5 * q23);
    double d23_0 = k0_23 * q23 * il23;
    const double v23 = k0_23;
    const double v24 = v22 + v23;
    const double v25 = v19 + v24;
    const double v26 = v14 * v25;
    // node 27: leaf op 2 at hp[15]
    const double il27 = ih[15], ip27 = ih[16];
    const double u27 = 3.14159265358979323846 * (l1 / h[16]);
    double s27, c27; gpb_sincos(u27, &s27, &c27);
    const double sine27 = s27 * s27;
    const double il2_27 = il27 * il27;
    const double k0_27 = exp((-2.0 * sine27) * il2_27);
    double d27_0 = k0_27 * (4.0 * sine27) * (il2_27 * il27);
    double d27_1 = k0_27 * (2.0 * 3.14159265358979323846 * l1 * (2.0 * s27 * c27)) * (il2_27 * (ip27 * ip27));
    const double v27 = k0_27;
    // node 28: leaf op 3 at hp[17]
    double k0_28 = 0.0;
    k0_28 += (xi[0] - h[17]) * (xj[0] - h[17]);
    double d28_0 = 2.0 * h[17] - xi[0] - xj[0];
    const double v28 = k0_28;
    const double v29 = v27 + v28;
    // node 30: leaf op 3 at hp[18]
    double k0_30 = 0.0;
    k0_30 += (xi[0] - h[18]) * (xj[0] - h[18]);
    double d30_0 = 2.0 * h[18] - xi[0] - xj[0];
    const double v30 = k0_30;
    const double v31 = v29 + v30;
    // node 32: leaf op 3 at hp[19]
    double k0_32 = 0.0;
    k0_32 += (xi[0] - h[19]) * (xj[0] - h[19]);
    double d32_0 = 2.0 * h[19] - xi[0] - xj[0];
    const double v32 = k0_32;
    // node 33: leaf op 3 at hp[20]
    double k0_33 = 0.0;
    k0_33 += (xi[0] - h[20]) * (xj[0] - h[20]);
    double d33_0 = 2.0 * h[20] - xi[0] - xj[0];
    const double v33 = k0_33;
    const double v34 = v32 * v33;
    // node 35: leaf op 1 at hp[21]
    const double il35 = ih[21];
    const double q35 = r2 * (il35 * il35);
    const double k0_35 = exp(-0.5 * q35);
    double d35_0 = k0_35 * q35 * il35;
    const double v35 = k0_35;
    const double v36 = v34 * v35;
    const double v37 = v31 * v36;
    // node 38: leaf op 3 at hp[22]
    double k0_38 = 0.0;
    k0_38 += (xi[0] - h[22]) * (xj[0] - h[22]);
    double d38_0 = 2.0 * h[22] - xi[0] - xj[0];
    const double v38 = k0_38;
    // node 39: leaf op 3 at hp[23]
    double k0_39 = 0.0;
    k0_39 += (xi[0] - h[23]) * (xj[0] - h[23]);
    double d39_0 = 2.0 * h[23] - xi[0] - xj[0];
    const double v39 = k0_39;
    const double v40 = v38 + v39;
    const double v41 = v37 * v40;
    const double v42 = v26 * v41;
    const double a42 = w;
    const double a26 = a42 * v41;
    const double a41 = a42 * v26;
    const double a37 = a41 * v40;
    const double a40 = a41 * v37;
    const double a38 = a40;
    const double a39 = a40;
    g[23] += a39 * d39_0;
    g[22] += a38 * d38_0;
    const double a31 = a37 * v36;
    const double a36 = a37 * v31;
    const double a34 = a36 * v35;
    const double a35 = a36 * v34;
    g[21] += a35 * d35_0;
    const double a32 = a34 * v33;
    const double a33 = a34 * v32;
    g[20] += a33 * d33_0;
    g[19] += a32 * d32_0;
    const double a29 = a31;
    const double a30 = a31;
    g[18] += a30 * d30_0;
    const double a27 = a29;
    const double a28 = a29;
    g[17] += a28 * d28_0;
    g[15] += a27 * d27_0;
    g[16] += a27 * d27_1;
    const double a14 = a26 * v25;
    const double a25 = a26 * v14;
    const double a19 = a25;
    const double a24 = a25;
    const double a22 = a24;
    const double a23 = a24;
    g[14] += a23 * d23_0;
    const double a20 = a22;
    const double a21 = a22;
    g[12] += a21 * d21_0;
    g[13] += a21 * d21_1;
    g[11] += a20 * d20_0;
    const double a17 = a19;
    const double a18 = a19;
    g[10] += a18 * d18_0;
    const double a15 = a17;
    const double a16 = a17;
    g[9] += a16 * d16_0;
    g[8] += a15 * d15_0;
    const double a10 = a14;
    const double a13 = a14;
    const double a11 = a13 * v12;
    const double a12 = a13 * v11;
    g[7] += a12 * d12_0;
    g[6] += a11 * d11_0;
    const double a4 = a10;
    const double a9 = a10;
    const double a7 = a9;
    const double a8 = a9;
    g[5] += a8 * d8_0;
    const double a5 = a7;
    const double a6 = a7;
    g[4] += a6 * d6_0;
    g[3] += a5 * d5_0;
    const double a2 = a4;
    const double a3 = a4;
    g[2] += a3 * d3_0;
    const double a0 = a2;
    const double a1 = a2;
    g[1] += a1 * d1_0;
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
