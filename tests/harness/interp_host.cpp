// TEST HARNESS ONLY: compiles csrc/program.cuh for the host so that the kernel-program interpreter (values and
// reverse-mode derivatives) can be compared with the oracle without a GPU.  Never linked into libgpb.
#include "../../gaussianprocessfundamentals_b200/csrc/program.cuh"

struct HostAcc {
  double* g;
  void operator()(int p, double v) const { g[p] += v; }
};

extern "C" {
// K[i*m + j] = k(X[i], X2[j]); if grad != null: grad[p] += sum_ij W[i*m+j] * dK_ij/dhp_p
int h_matrix(const int32_t* code, int n_ops, int dim, int cp_mode, const double* X, int n, const double* X2, int m,
             const double* hp, double* K, const double* W, double* grad) {
  GpbPair p;
  double ihp[GPB_MAX_HP + 1];
  int n_hp = 0;
  for (int pc = 0; pc < n_ops; ++pc) {
    const int32_t* w = code + pc * GPB_OP_WORDS;
    int top = w[1] + (w[0] < GPB_OP_ADD2 ? gpb_leaf_nhp(w[0], w[2], dim) : (w[0] == GPB_OP_CPW ? w[3] - 1 : 0));
    if (top > n_hp) n_hp = top;
  }
  for (int i = 0; i < n_hp; ++i) ihp[i] = 1.0 / hp[i];
  p.dim = dim; p.hp = hp; p.ihp = ihp; p.cp_mode = cp_mode;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) {
      p.xi = X + (size_t)i * dim; p.xj = X2 + (size_t)j * dim; p.gi = i; p.gj = j;
      double v = gpb_eval(code, n_ops, p);
      if (grad) {
        HostAcc acc{grad};
        double v2 = gpb_eval_grad(code, n_ops, p, W[(size_t)i * m + j], acc);
        if (v2 != v) return 1;
      }
      K[(size_t)i * m + j] = v;
    }
  return 0;
}
}
